/*
 * results.c -- what the reference's driver does with the probabilities: softmax
 * (ViT_seq.c:304-324), argmax + result file (Main.c:62-72) and the golden-file check
 * (comparator.c:23-80).  Plain C, no CUDA.
 */
#include "vit_host.h"

#include <math.h>
#include <stdio.h>
#include <string.h>

void vit_softmax(const float* logits, float* probs, int length) {
    float max_val = logits[0];
    for (int i = 1; i < length; ++i)
        if (logits[i] > max_val) max_val = logits[i];
    float sum_exp = 0.0f;
    for (int i = 0; i < length; ++i) {
        probs[i] = expf(logits[i] - max_val);
        sum_exp += probs[i];
    }
    for (int i = 0; i < length; ++i) probs[i] /= sum_exp;
}

int vit_argmax(const float* v, int length) {
    int best = 0;
    for (int j = 1; j < length; ++j)
        if (v[j] > v[best]) best = j;
    return best;
}

int write_results(const char* filename, float* const* prb, int n) {
    FILE* fp = fopen(filename, "w");
    if (!fp) {
        fprintf(stderr, "Error: cannot open %s for writing\n", filename);
        return -1;
    }
    for (int i = 0; i < n; ++i) {
        const int pred = vit_argmax(prb[i], VIT_NUM_CLASSES);
        fprintf(fp, "[%d] label: %d / prob: %.6f\n", i, pred, prb[i][pred]);
    }
    fclose(fp);
    return 0;
}

static int parse_result_line(const char* line, int* label, float* prob) {
    return sscanf(line, "[%*d] label: %d / prob: %f)", label, prob);
}

int comparator_files(const char* result_path, const char* answer_path, int image_count) {
    FILE* fr = fopen(result_path, "r");
    if (!fr) {
        fprintf(stderr, "Error: Cannot open %s\n", result_path);
        return 1;
    }
    FILE* fa = fopen(answer_path, "r");
    if (!fa) {
        fprintf(stderr, "Error: Cannot open %s\n", answer_path);
        fclose(fr);
        return 1;
    }
    char lr[1024], la[1024];
    int errors = 0;
    for (int line = 0; line < image_count; ++line) {
        if (!fgets(lr, sizeof(lr), fr) || !fgets(la, sizeof(la), fa)) {
            fprintf(stderr, "Line %d: not enough lines\n", line);
            ++errors;
            break;
        }
        int label_r, label_a;
        float prob_r, prob_a;
        if (parse_result_line(lr, &label_r, &prob_r) != 2 || parse_result_line(la, &label_a, &prob_a) != 2) {
            fprintf(stderr, "Line %d: parse error\n", line);
            ++errors;
            continue;
        }
        if (label_r != label_a) {
            fprintf(stderr, "Line %d: Label mismatch (Result: %d, Answer: %d)\n", line, label_r, label_a);
            ++errors;
        }
        if (fabs(prob_r - prob_a) > 0.01f) {
            fprintf(stderr, "Line %d: Probability mismatch (Result: %.6f, Answer: %.6f)\n", line, prob_r, prob_a);
            ++errors;
        }
    }
    fclose(fr);
    fclose(fa);
    return errors;
}
