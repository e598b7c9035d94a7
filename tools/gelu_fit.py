#!/usr/bin/env python
"""Minimax fit of x * Phi(x) by x * sigmoid(x (a + b x^2 + c x^4 + d x^6)) (the GELU of the mlp_0 epilogue, csrc/ptx.cuh gelu_erf2):
Nelder-Mead on the maximum absolute error over [-8.5, 8.5] with a penalty that keeps the polynomial monotone out to |x| = 300."""
import numpy as np
from scipy.special import erf, ndtr, log_ndtr
from scipy.optimize import minimize
x=np.linspace(-8.5,8.5,34001)
gelu=x*ndtr(x)
def poly(p,x):
    x2=x*x
    acc=p[-1]
    for c in p[-2::-1]: acc=acc*x2+c
    return x*acc
def model(p,x):
    with np.errstate(over='ignore'):
        return x/(1+np.exp(-poly(p,x)))
def obj(p):
    e=np.max(np.abs(model(p,x)-gelu))
    # monotonic far range: p(x) must stay >= 30 for x in [9, 300]
    xf=np.array([9,10,12,16,24,40,80,160,300.0])
    pen=np.sum(np.maximum(0,30-poly(p,xf)))
    return e+1e-3*pen
# initial via least squares on g(t)=logit(Phi)/x
xs=np.linspace(0.05,6,400)
g=(log_ndtr(xs)-log_ndtr(-xs))/xs
for n in (3,4,5):
    A=np.stack([xs**(2*k) for k in range(n)],1)
    p0=np.linalg.lstsq(A,g,rcond=None)[0]
    best=None
    for trial in range(6):
        p=p0*(1+0.02*np.random.default_rng(trial).standard_normal(n))
        r=minimize(obj,p,method='Nelder-Mead',options={'xatol':1e-13,'fatol':1e-15,'maxiter':60000,'maxfev':60000})
        for _ in range(4):
            r=minimize(obj,r.x,method='Nelder-Mead',options={'xatol':1e-14,'fatol':1e-16,'maxiter':60000,'maxfev':60000})
        if best is None or r.fun<best.fun: best=r
    xx=np.linspace(-300,300,600001)
    e_all=np.max(np.abs(model(best.x,xx)-xx*ndtr(xx)))
    print(n,[float(v) for v in best.x],"err[-8.5,8.5] %.3e  err[-300,300] %.3e"%(np.max(np.abs(model(best.x,x)-gelu)),e_all))
