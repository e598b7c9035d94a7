/* Empty stand-in for the OpenCL header: the reference's ViT_seq.c includes <CL/cl.h>
 * (ViT_seq.c:8) but uses nothing from it.  Used only when building oracle/_ref. */
