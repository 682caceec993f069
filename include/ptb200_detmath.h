/* ptb200_detmath.h — deterministic sin/cos for the FP64 validation mode.
 *
 * Why: the reference draws scatter directions with libm sin/cos (src/smallpt.cpp:347,359).
 * glibc and CUDA libm differ in the last bit for a fraction of arguments, and on the
 * no-epsilon rectangle scene one flipped bit desynchronises the rest of the row's erand48
 * stream.  This header is the shared replacement (oracle patch "P6"):
 * only +, -, *, / and int<->double conversion, evaluated in a fixed order, so that
 * gcc -ffp-contract=off and nvcc -fmad=false produce bit-identical results.
 *
 * Domain: a in [0, 2*pi] (r1 = 2*M_PI*erand48).  Accuracy ~1 ulp (tested against libm).
 * Cody-Waite two-constant reduction by pi/2, then the classic degree-13/14 minimax kernels
 * on [-pi/4, pi/4].
 */
#ifndef PTB200_DETMATH_H
#define PTB200_DETMATH_H

#if defined(__CUDACC__)
#define PT_HD __host__ __device__ __forceinline__
#else
#define PT_HD static inline
#endif

PT_HD void pt_det_sincos(double a, double *s_out, double *c_out)
{
    const double two_over_pi = 6.36619772367581382433e-01;
    const double pio2_hi = 1.57079632673412561417e+00;   /* first 33 bits of pi/2 */
    const double pio2_lo = 6.07710050650619224932e-11;   /* pi/2 - pio2_hi        */
    int    k = (int)(a * two_over_pi + 0.5);
    double kd = (double)k;
    double x = (a - kd * pio2_hi) - kd * pio2_lo;         /* |x| <= pi/4 (+ tiny)   */
    double z = x * x;
    /* sin kernel */
    double rs = 8.33333333332248946124e-03 + z * (-1.98412698298579493134e-04 + z * (2.75573137070700676789e-06
              + z * (-2.50507602534068634195e-08 + z * 1.58969099521155010221e-10)));
    double sn = x + (z * x) * (-1.66666666666666324348e-01 + z * rs);
    /* cos kernel */
    double rc = z * (4.16666666666666019037e-02 + z * (-1.38888888888741095749e-03 + z * (2.48015872894767294178e-05
              + z * (-2.75573143513906633035e-07 + z * (2.08757232129817482790e-09 + z * -1.13596475577881948265e-11)))));
    double cs = 1.0 - (0.5 * z - z * rc);
    switch (k & 3) {
    case 0:  *s_out = sn;  *c_out = cs;  break;
    case 1:  *s_out = cs;  *c_out = -sn; break;
    case 2:  *s_out = -sn; *c_out = -cs; break;
    default: *s_out = -cs; *c_out = sn;  break;
    }
}

PT_HD double pt_det_sin(double a) { double s, c; pt_det_sincos(a, &s, &c); return s; }
PT_HD double pt_det_cos(double a) { double s, c; pt_det_sincos(a, &s, &c); return c; }

#endif /* PTB200_DETMATH_H */
