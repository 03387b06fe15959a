// pbc.cuh -- device-side periodic-boundary geometry.
//
// Two families:
//   *_exact : the reference's arithmetic, operation by operation, with explicit round-to-nearest
//             intrinsics so that nvcc never contracts a*b+c into an FMA.  Results are bit-identical
//             to the CPU oracle (oracle/cmdlmc_oracle.c, compiled with -ffp-contract=off) and
//             agree with the reference build (which is -ffast-math) to ~1e-15 relative.
//   *_fast  : FMA arithmetic with a pruned image set, used only as a conservative FILTER in the
//             all-pairs kernels; every pair that survives the filter is re-evaluated exactly.
#pragma once
#include "common.cuh"

// ---- A1: numpyatom.pyx:33-42 (diff_ptr): repeated +-L while outside [-L/2, L/2] (strict) ------
__device__ __forceinline__ double wrap_ortho_exact(double d, double L, double hL)
{
    // Safety valve (documented divergence): the reference loops |d|/L times; beyond 64 box
    // lengths we first remove whole boxes with one rounding and then finish with the exact loop.
    if (fabs(d) > 64.0 * L) d = __dadd_rn(d, -__dmul_rn(L, rint(d / L)));
    if (!(fabs(d) <= 128.0 * L)) return d;  // inf / nan: the reference would never return
    while (d < -hL) d = __dadd_rn(d, L);
    while (d > hL) d = __dadd_rn(d, -L);
    return d;
}

__device__ __forceinline__ void diff_ortho_exact(const BoxParams &bx, const double a[3],
                                                 const double b[3], double d[3])
{
#pragma unroll
    for (int i = 0; i < 3; i++) d[i] = wrap_ortho_exact(__dadd_rn(b[i], -a[i]), bx.L[i], bx.hL[i]);
}

// dx*dx + dy*dy + dz*dz in the reference's order, no FMA (numpyatom.pyx:179)
__device__ __forceinline__ double norm2_exact(const double d[3])
{
    return __dadd_rn(__dadd_rn(__dmul_rn(d[0], d[0]), __dmul_rn(d[1], d[1])), __dmul_rn(d[2], d[2]));
}

// math_helper.pyx:50-60 (matrix_mult_ptr): r_i = ((0 + m_i0 v0) + m_i1 v1) + m_i2 v2
__device__ __forceinline__ void matvec3_exact(const double m[9], double v[3])
{
    double r[3];
#pragma unroll
    for (int i = 0; i < 3; i++)
        r[i] = __dadd_rn(__dadd_rn(__dadd_rn(0.0, __dmul_rn(m[3 * i], v[0])),
                                   __dmul_rn(m[3 * i + 1], v[1])),
                         __dmul_rn(m[3 * i + 2], v[2]));
    v[0] = r[0]; v[1] = r[1]; v[2] = r[2];
}

// x - round(x) with C99 round (half away from zero), bit for bit, without libm's round():
// r = rint(x) by the 1.5*2^52 trick (exact for |x| < 2^51), x - r is exact, and round() differs
// from rint() only on ties that rint resolved towards zero.
__device__ __forceinline__ double sub_round_exact(double x)
{
    if (!(fabs(x) < 2251799813685248.0)) return __dadd_rn(x, -round(x));   // cold: |x| >= 2^51, nan
    const double magic = 6755399441055744.0;
    const double r = __dadd_rn(__dadd_rn(x, magic), -magic);
    double w = __dadd_rn(x, -r);
    if (w == 0.5 && x > 0.0) w = -0.5;
    if (w == -0.5 && x < 0.0) w = 0.5;
    return w;
}

// The three components at once: the straight-line path (rint by the magic constant, exact
// subtraction) plus ONE test on the high words for everything that needs sub_round_exact's care:
// |x| >= 2^51 / nan / inf, or a tie |x - rint(x)| == 1/2 (|w| <= 1/2 always, so >= suffices).
__device__ __forceinline__ void sub_round3_exact(double s[3])
{
    const double magic = 6755399441055744.0;
    double w[3];
    int hs = 0, hw = 0;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        w[i] = __dadd_rn(s[i], -__dadd_rn(__dadd_rn(s[i], magic), -magic));
        hs = max(hs, __double2hiint(s[i]) & 0x7fffffff);
        hw = max(hw, __double2hiint(w[i]) & 0x7fffffff);
    }
    if (hs >= 0x43200000 || hw >= 0x3fe00000) {   // cold
#pragma unroll
        for (int i = 0; i < 3; i++) s[i] = sub_round_exact(s[i]);
        return;
    }
    s[0] = w[0]; s[1] = w[1]; s[2] = w[2];
}

// ---- A3: numpyatom.pyx:61-74 (diff_ptr_nonortho); round = C99 round, half away from zero ------
__device__ __forceinline__ void diff_general_exact(const BoxParams &bx, const double a[3],
                                                   const double b[3], double d[3])
{
#pragma unroll
    for (int i = 0; i < 3; i++) d[i] = __dadd_rn(b[i], -a[i]);
    matvec3_exact(bx.hinv, d);
    sub_round3_exact(d);
    matvec3_exact(bx.h, d);
}

// The same vector for callers that only take its NORM: matrix_mult_ptr's leading `0 +` is left
// out.  0 + x differs from x only for x = -0, and a flipped zero sign can only flip the sign of
// zero-valued components further down -- the squared length is bit-identical.
__device__ __forceinline__ void matvec3_norm_exact(const double m[9], double v[3])
{
    double r[3];
#pragma unroll
    for (int i = 0; i < 3; i++)
        r[i] = __dadd_rn(__dadd_rn(__dmul_rn(m[3 * i], v[0]), __dmul_rn(m[3 * i + 1], v[1])),
                         __dmul_rn(m[3 * i + 2], v[2]));
    v[0] = r[0]; v[1] = r[1]; v[2] = r[2];
}

__device__ __forceinline__ void diff_general_norm_exact(const BoxParams &bx, const double a[3],
                                                        const double b[3], double d[3])
{
#pragma unroll
    for (int i = 0; i < 3; i++) d[i] = __dadd_rn(b[i], -a[i]);
    matvec3_norm_exact(bx.hinv, d);
    sub_round3_exact(d);
    matvec3_norm_exact(bx.h, d);
}

// The same again for cell matrices with structural zeros (BoxParams::sparse, set by
// cmd_box_sparsity): a cell given in the usual convention -- a along x, b in the xy plane -- has an
// upper-triangular h (columns = cell vectors) and so has its inverse; a monoclinic cell with unique
// axis b keeps four entries.  A product with an exact zero is +-0 and x + (+-0) == x up to the
// sign of a zero result, which the squared length cannot see (same argument as above), so leaving
// those terms out is bit-identical for finite coordinates.
//   SP 1: m[3] = m[6] = m[7] = 0          SP 2: additionally m[1] = m[5] = 0
template <int SP>
__device__ __forceinline__ void matvec3_norm_sp(const double m[9], double v[3])
{
    if (SP == 0) { matvec3_norm_exact(m, v); return; }
    double r0;
    if (SP == 1) r0 = __dadd_rn(__dadd_rn(__dmul_rn(m[0], v[0]), __dmul_rn(m[1], v[1])), __dmul_rn(m[2], v[2]));
    else r0 = __dadd_rn(__dmul_rn(m[0], v[0]), __dmul_rn(m[2], v[2]));
    const double r1 = SP == 1 ? __dadd_rn(__dmul_rn(m[4], v[1]), __dmul_rn(m[5], v[2])) : __dmul_rn(m[4], v[1]);
    const double r2 = __dmul_rn(m[8], v[2]);
    v[0] = r0; v[1] = r1; v[2] = r2;
}

template <int SP>
__device__ __forceinline__ void diff_general_norm_sp(const BoxParams &bx, const double a[3],
                                                     const double b[3], double d[3])
{
#pragma unroll
    for (int i = 0; i < 3; i++) d[i] = __dadd_rn(b[i], -a[i]);
    matvec3_norm_sp<SP>(bx.hinv, d);
    sub_round3_exact(d);
    matvec3_norm_sp<SP>(bx.h, d);
}

// host: structural zeros of BOTH h and hinv (0, 1 or 2 as above; 0 for orthorhombic boxes, which
// never take this path)
static inline int cmd_box_sparsity(const BoxParams &p)
{
    if (p.kind == 0) return 0;
    auto z = [&](int k) { return p.h[k] == 0.0 && p.hinv[k] == 0.0; };
    if (!(z(3) && z(6) && z(7))) return 0;
    return z(1) && z(5) ? 2 : 1;
}

// ---- A4: numpyatom.pyx:101-123: min over the 27 images of the wrapped vector, squared ---------
__device__ __forceinline__ double min_image_norm2_exact(const BoxParams &bx, const double d[3])
{
    double mind = 1e6;
#pragma unroll 1
    for (int i = -1; i < 2; i++) {
        double u[3];
#pragma unroll
        for (int c = 0; c < 3; c++) u[c] = __dadd_rn(d[c], i * bx.h[3 * c]);
#pragma unroll
        for (int j = -1; j < 2; j++) {
            double w[3];
#pragma unroll
            for (int c = 0; c < 3; c++) w[c] = __dadd_rn(u[c], j * bx.h[3 * c + 1]);
#pragma unroll
            for (int k = -1; k < 2; k++) {
                double v[3];
#pragma unroll
                for (int c = 0; c < 3; c++) v[c] = __dadd_rn(w[c], k * bx.h[3 * c + 2]);
                double n2 = norm2_exact(v);
                if (n2 < mind) mind = n2;
            }
        }
    }
    return mind;
}

// squared reference length of (b - a), before sqrt and before the water conversion
__device__ __forceinline__ double length2_exact(const BoxParams &bx, const double a[3],
                                                const double b[3])
{
    double d[3];
    if (bx.kind == 0) {
        diff_ortho_exact(bx, a, b, d);
        return norm2_exact(d);
    }
    diff_general_exact(bx, a, b, d);
    return min_image_norm2_exact(bx, d);
}

// PBCHelper.pyx:318-324, 342-351 (convert_distance)
__device__ __forceinline__ double convert_distance(const BoxParams &bx, double d)
{
    if (bx.conv == CMD_CONV_NONE) return d;
    if (bx.conv_par[3] < d && d < bx.conv_par[4]) {
        if (bx.conv == CMD_CONV_LINEAR) return __dadd_rn(__dmul_rn(bx.conv_par[0], d), bx.conv_par[1]);
        if (d < bx.conv_par[2]) return bx.conv_par[1];
        return __dadd_rn(__dmul_rn(bx.conv_par[0], __dadd_rn(d, -bx.conv_par[2])), bx.conv_par[1]);
    }
    return d;
}

// AtomBox.length_ptr dispatch (PBCHelper.pyx:228-232, 262-268, 300-303)
__device__ __forceinline__ double length_exact(const BoxParams &bx, const double a[3],
                                               const double b[3])
{
    return convert_distance(bx, sqrt(length2_exact(bx, a, b)));
}

// AtomBox.distance_vector dispatch (PBCHelper.pyx:234-235, 270-271)
__device__ __forceinline__ void distance_exact(const BoxParams &bx, const double a[3],
                                               const double b[3], double d[3])
{
    if (bx.kind == 0) diff_ortho_exact(bx, a, b, d);
    else diff_general_exact(bx, a, b, d);
}

// math_helper.pyx:16-23 (dot_product_ptr): ((0 + a0 b0) + a1 b1) + a2 b2
__device__ __forceinline__ double dot3_exact(const double a[3], const double b[3])
{
    return __dadd_rn(__dadd_rn(__dadd_rn(0.0, __dmul_rn(a[0], b[0])), __dmul_rn(a[1], b[1])),
                     __dmul_rn(a[2], b[2]));
}

// ---- A5: AtomBox.angle_ptr: angle at p2 between p1 and p3 -------------------------------------
// ortho (numpyatom.pyx:244-264, wraps '>' first then '<'); general (numpyatom.pyx:280-291,
// fractional wrap only).
__device__ __forceinline__ double angle_exact(const BoxParams &bx, const double p1[3],
                                              const double p2[3], const double p3[3])
{
    double v1[3], v2[3];
    if (bx.kind == 0) {
#pragma unroll
        for (int i = 0; i < 3; i++) {
            // same fixed point as diff_ptr: the two loops commute for a finite box
            v1[i] = wrap_ortho_exact(__dadd_rn(p1[i], -p2[i]), bx.L[i], bx.hL[i]);
            v2[i] = wrap_ortho_exact(__dadd_rn(p3[i], -p2[i]), bx.L[i], bx.hL[i]);
        }
    } else {
        diff_general_exact(bx, p2, p1, v1);
        diff_general_exact(bx, p2, p3, v2);
    }
    return acos(dot3_exact(v1, v2) / sqrt(dot3_exact(v1, v1)) / sqrt(dot3_exact(v2, v2)));
}

// ---- FAST filter: a value within ~1e-14 relative of the reference length^2 ---------------------
// rint via the 1.5*2^52 trick (2 DADD on the FP64 pipe instead of a conversion instruction);
// valid for |x| < 2^51, which holds for fractional coordinates of any sane trajectory.
__device__ __forceinline__ double rint_magic(double x)
{
    const double magic = 6755399441055744.0;
    return __dadd_rn(__dadd_rn(x, magic), -magic);
}

__device__ __forceinline__ double length2_fast(const BoxParams &bx, double dx, double dy, double dz)
{
    if (bx.kind == 0) {
        dx -= bx.L[0] * rint_magic(dx * (1.0 / bx.L[0]));
        dy -= bx.L[1] * rint_magic(dy * (1.0 / bx.L[1]));
        dz -= bx.L[2] * rint_magic(dz * (1.0 / bx.L[2]));
        return fma(dz, dz, fma(dy, dy, dx * dx));
    }
    double s0 = fma(bx.hinv[2], dz, fma(bx.hinv[1], dy, bx.hinv[0] * dx));
    double s1 = fma(bx.hinv[5], dz, fma(bx.hinv[4], dy, bx.hinv[3] * dx));
    double s2 = fma(bx.hinv[8], dz, fma(bx.hinv[7], dy, bx.hinv[6] * dx));
    s0 -= rint_magic(s0);
    s1 -= rint_magic(s1);
    s2 -= rint_magic(s2);
    double cx = fma(bx.h[2], s2, fma(bx.h[1], s1, bx.h[0] * s0));
    double cy = fma(bx.h[5], s2, fma(bx.h[4], s1, bx.h[3] * s0));
    double cz = fma(bx.h[8], s2, fma(bx.h[7], s1, bx.h[6] * s0));
    double best = fma(cz, cz, fma(cy, cy, cx * cx));
    for (int m = 0; m < bx.n_img; m++) {
        double vx = cx + bx.img[m][0], vy = cy + bx.img[m][1], vz = cz + bx.img[m][2];
        best = fmin(best, fma(vz, vz, fma(vy, vy, vx * vx)));
    }
    return best;
}

// ---- jump rates (A9 / A9') ---------------------------------------------------------------------
struct RateParams {
    int kind;
    double par[CMD_RATE_NPAR];
};

#define CMD_KB_EV 8.617333262e-5

// a / den for den in [1, 1e305): reciprocal seed + two Newton steps + one residual correction
// (<= 1 ulp; the reference divides in IEEE, the rate gate is 1e-10 relative)
__device__ __forceinline__ double div_fast(double a, double den)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(den));
    double e = fma(-den, y, 1.0);
    y = fma(y, e, y);
    e = fma(-den, y, 1.0);
    y = fma(y, e, y);
    const double q = a * y;
    const double r = fma(-den, q, a);
    return fma(r, y, q);
}

// exp(z) for z in [-700, 700]: k = rint(z log2 e), r = z - k ln 2 (two-term Cody-Waite),
// Taylor polynomial of degree 13 on |r| <= 0.347 (truncation 4e-18), scaled by 2^k through the
// exponent field.  1.3 ulp measured against long double; no special cases (the callers clamp).
// The coefficients live in constant memory: an FP64 instruction takes a constant-bank operand
// directly, a 64-bit immediate costs two extra moves per use.
static __constant__ double EXP_COEF[16] = {
    1.4426950408889634, -6.93147180369123816490e-01, -1.90821492927058770002e-10,
    1.6059043836821613e-10, 2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07,
    2.7557319223985893e-06, 2.48015873015873e-05, 1.984126984126984e-04, 1.388888888888889e-03,
    8.333333333333333e-03, 4.1666666666666664e-02, 1.6666666666666666e-01, 0.5, 6755399441055744.0};

__device__ __forceinline__ double exp_core(double z)
{
    const double magic = EXP_COEF[15];
    const double t = fma(z, EXP_COEF[0], magic);
    const int k = __double2loint(t);
    const double kf = t - magic;
    double r = fma(kf, EXP_COEF[1], z);
    r = fma(kf, EXP_COEF[2], r);
    double p = EXP_COEF[3];                       // 1/13!
#pragma unroll
    for (int i = 4; i <= 14; i++) p = fma(p, r, EXP_COEF[i]);   // 1/12! ... 1/2!
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return p * __hiloint2double((k + 1023) << 20, 0);
}

// Fermi rate a / (1 + exp((x - b) / c)) (jumprate_generators.py:33-34): the division by c is a
// multiplication by the correctly rounded 1/c (loop invariant), exp is exp_core, the outer
// division div_fast.  One implementation for every kernel, so that lists built by different
// kernels agree bit for bit; a few ulp from NumPy's value (the gate is 1e-10 relative).
__device__ __forceinline__ double fermi_eval(const RateParams &r, double x)
{
    const double inv_c = __drcp_rn(r.par[2]);
    const double z = (x - r.par[1]) * inv_c;
    if (!(z < 700.0)) return r.par[0] / (1.0 + exp((x - r.par[1]) / r.par[2]));   // cold
    return div_fast(r.par[0], 1.0 + exp_core(fmax(z, -700.0)));
}

__device__ __forceinline__ double rate_eval(const RateParams &r, double x, double theta)
{
    switch (r.kind) {
    case CMD_RATE_FERMI:
        return fermi_eval(r, x);
    case CMD_RATE_FERMI_ANGLE:  // jumprate_generators.py:42-43
        return theta < r.par[3] ? 0.0 : fermi_eval(r, x);
    case CMD_RATE_AE: {  // IO/config_parser.py:334-342 (specification text; parity unpinned)
        // E = a u / sqrt(b + 1/u^2) = a u^2 / sqrt(b u^2 + 1) for u > 0: one reciprocal square root
        // instead of a division, a square root and another division; exp through exp_core.  A few
        // ulp from the restatement's libm value (the gate is 1e-10 relative).
        const double u = x - r.par[3];
        if (!(u > 0)) return r.par[0];
        const double u2 = u * u;
        const double e = r.par[1] * u2 * rsqrt(fma(r.par[2], u2, 1.0));
        const double z = -e * __drcp_rn(CMD_KB_EV * r.par[4]);
        if (!(z > -700.0) || !(u2 < 1e300))   // cold: underflowing rate, overflowing u^2, nan
            return r.par[0] * exp(-(r.par[1] * u / sqrt(r.par[2] + 1.0 / (u * u))) / (CMD_KB_EV * r.par[4]));
        return r.par[0] * exp_core(z);
    }
    case CMD_RATE_EXP:  // IO/config_parser.py:344-345
        return r.par[0] * exp(r.par[1] * x);
    }
    return 0.0;
}

// two independent evaluations behind ONE dispatch, so their FP64 chains can interleave
__device__ __forceinline__ void rate_eval2(const RateParams &r, double x0, double x1, double out[2])
{
    if (r.kind == CMD_RATE_FERMI) {
        out[0] = fermi_eval(r, x0);
        out[1] = fermi_eval(r, x1);
    } else {
        out[0] = rate_eval(r, x0, 0.0);
        out[1] = rate_eval(r, x1, 0.0);
    }
}
