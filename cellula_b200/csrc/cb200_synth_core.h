// cb200_synth_core.h -- THE synthetic-data generator of this repository (bench / test infrastructure,
// not part of the reference path).  One specification, compiled twice from this header:
//   * for the device (cb200_synth.cu: the bench generates 1.3 M - 4 M cells in HBM), and
//   * for the host  (cb200_host.cu: cb200_synth_host / cb200_synth_embedding_host -> the parity tests, the
//     CPU baseline legs of bench.py and scripts/write_dataset.py, which writes the language-neutral on-disk
//     form -- p.i64 / i.i32 / x.f64 / manifest.json -- that baseline/run_reference.R reads on an R host).
// Both compilations produce the SAME matrix bit for bit (tests/test_gpu_synth.py): every draw is a pure
// function of (seed, global cell index, gene), and everything is computed with integer operations and
// individually rounded IEEE-754 add / multiply / divide / sqrt only (no FMA contraction, no libm / CUDA math
// functions whose last bit differs between host and device); exp, log and the normal quantile are restated
// below in those operations.
//
// Model: negative-binomial counts with a planted low-rank structure,
//   log lambda_gc = base_g + off + log depth_c + sum_z W_g,z * Z_{f(g,z)},c
// F = 64 latent factors of geometrically decaying strength (0.93^f), every gene loading on 4 of them,
// Z_c = centroid[cluster(c)] + 0.3 N(0,1) over 40 clusters (continuous noise: no duplicate cells),
// depth_c ~ LogNormal(0, 0.35^2), counts ~ NB(mean lambda, size 10); `off` calibrates the mean number of
// non-zero genes per cell.  An all-zero cell gets one count (logNormCounts would stop on it).
#pragma once

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define SY_HD __host__ __device__ __forceinline__
#else
#define SY_HD static inline
#endif

#define SY_F 64
#define SY_NZ 4
#define SY_CLUSTERS 40
#define SY_SIZE 10.0f      // NB size parameter r = 1 / dispersion (sy_p0 relies on r == 10)
#define SY_LL_MAX 12.0f
#define SY_EMB_CLUSTERS 40

// ---- individually rounded arithmetic -------------------------------------------------------------------
SY_HD float sy_fmul(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
SY_HD float sy_fadd(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
SY_HD float sy_fdiv(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
SY_HD float sy_fsqrt(float a)
{
#ifdef __CUDA_ARCH__
    return __fsqrt_rn(a);
#else
    return __builtin_sqrtf(a);
#endif
}
SY_HD double sy_dmul(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
SY_HD double sy_dadd(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
SY_HD double sy_ddiv(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}
SY_HD double sy_dsqrt(double a)
{
#ifdef __CUDA_ARCH__
    return __dsqrt_rn(a);
#else
    return __builtin_sqrt(a);
#endif
}
SY_HD float sy_bits_to_float(uint32_t b)
{
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}
SY_HD uint64_t sy_double_to_bits(double d)
{
#ifdef __CUDA_ARCH__
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t b;
    memcpy(&b, &d, 8);
    return b;
#endif
}
SY_HD double sy_bits_to_double(uint64_t b)
{
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)b);
#else
    double d;
    memcpy(&d, &b, 8);
    return d;
#endif
}

// ---- counter-based randomness ----------------------------------------------------------------------------
SY_HD uint64_t sy_mix(uint64_t z)  // splitmix64 finaliser
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
SY_HD uint64_t sy_hash(uint64_t seed, uint64_t a, uint64_t b)
{
    return sy_mix(sy_mix(seed + 0x9E3779B97F4A7C15ULL * (a + 1)) ^ (0xD1B54A32D192ED03ULL * (b + 1)));
}
SY_HD float sy_unif(uint64_t h)  // (0, 1), 23 bits: (m + 0.5) / 2^23 is exact in fp32
{
    return sy_fmul(sy_fadd((float)(uint32_t)(h >> 41), 0.5f), 1.0f / 8388608.0f);
}
SY_HD double sy_unif_d(uint64_t h)  // (0, 1), 52 bits, exact in fp64
{
    return sy_dmul(sy_dadd((double)(h >> 12), 0.5), 1.0 / 4503599627370496.0);
}

// ---- exp (fp32), log (fp64), normal quantile (fp64) in rounded operations only -----------------------------
SY_HD float sy_expf(float x)  // x in [-80, 80]; relative error ~2e-7
{
    if (x < -80.f)
        x = -80.f;
    const float t = sy_fmul(x, 1.44269504f);
    const int n = (int)sy_fadd(t, t < 0.f ? -0.5f : 0.5f);
    const float fn = (float)n;
    float r = sy_fadd(x, -sy_fmul(fn, 0.693145752f));   // ln2 high part: fn * hi is exact
    r = sy_fadd(r, -sy_fmul(fn, 1.42860677e-06f));      // ln2 low part
    float p = 1.0f / 5040.0f;                           // Taylor, |r| <= 0.347: remainder < 1e-8
    p = sy_fadd(sy_fmul(p, r), 1.0f / 720.0f);
    p = sy_fadd(sy_fmul(p, r), 1.0f / 120.0f);
    p = sy_fadd(sy_fmul(p, r), 1.0f / 24.0f);
    p = sy_fadd(sy_fmul(p, r), 1.0f / 6.0f);
    p = sy_fadd(sy_fmul(p, r), 0.5f);
    p = sy_fadd(sy_fmul(p, r), 1.0f);
    p = sy_fadd(sy_fmul(p, r), 1.0f);
    return sy_fmul(p, sy_bits_to_float((uint32_t)(n + 127) << 23));  // n in [-116, 116]
}

SY_HD double sy_log(double x)  // x > 0, normal; atanh series with fdlibm's coefficients
{
    const uint64_t b = sy_double_to_bits(x);
    int e = (int)((b >> 52) & 0x7ff) - 1023;
    double m = sy_bits_to_double((b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);  // [1, 2)
    if (m > 1.4142135623730951) {
        m = sy_dmul(m, 0.5);
        e += 1;
    }
    const double f = sy_dadd(m, -1.0);
    const double s = sy_ddiv(f, sy_dadd(2.0, f));
    const double z = sy_dmul(s, s);
    double R = 1.479819860511658591e-01;
    R = sy_dadd(sy_dmul(R, z), 1.531383769920937332e-01);
    R = sy_dadd(sy_dmul(R, z), 1.818357216161805012e-01);
    R = sy_dadd(sy_dmul(R, z), 2.222219843214978396e-01);
    R = sy_dadd(sy_dmul(R, z), 2.857142874366239149e-01);
    R = sy_dadd(sy_dmul(R, z), 3.999999999940941908e-01);
    R = sy_dadd(sy_dmul(R, z), 6.666666666666735130e-01);
    R = sy_dmul(R, z);
    const double l1p = sy_dadd(sy_dmul(2.0, s), sy_dmul(s, R));
    return sy_dadd(sy_dmul((double)e, 0.6931471805599453), l1p);
}

SY_HD double sy_ndtri(double p)  // standard normal quantile (Acklam's rational approximation, |rel err| ~ 1e-9)
{
    const double a0 = -3.969683028665376e+01, a1 = 2.209460984245205e+02, a2 = -2.759285104469687e+02,
                 a3 = 1.383577518672690e+02, a4 = -3.066479806614716e+01, a5 = 2.506628277459239e+00;
    const double b0 = -5.447609879822406e+01, b1 = 1.615858368580409e+02, b2 = -1.556989798598866e+02,
                 b3 = 6.680131188771972e+01, b4 = -1.328068155288572e+01;
    const double c0 = -7.784894002430293e-03, c1 = -3.223964580411365e-01, c2 = -2.400758277161838e+00,
                 c3 = -2.549732539343734e+00, c4 = 4.374664141464968e+00, c5 = 2.938163982698783e+00;
    const double d0 = 7.784695709041462e-03, d1 = 3.224671290700398e-01, d2 = 2.445134137142996e+00,
                 d3 = 3.754408661907416e+00;
    const double plow = 0.02425;
    if (p < plow || p > 1.0 - plow) {
        const bool upper = p > 0.5;
        const double pp = upper ? sy_dadd(1.0, -p) : p;
        const double q = sy_dsqrt(sy_dmul(-2.0, sy_log(pp)));
        double num = c0;
        num = sy_dadd(sy_dmul(num, q), c1);
        num = sy_dadd(sy_dmul(num, q), c2);
        num = sy_dadd(sy_dmul(num, q), c3);
        num = sy_dadd(sy_dmul(num, q), c4);
        num = sy_dadd(sy_dmul(num, q), c5);
        double den = d0;
        den = sy_dadd(sy_dmul(den, q), d1);
        den = sy_dadd(sy_dmul(den, q), d2);
        den = sy_dadd(sy_dmul(den, q), d3);
        den = sy_dadd(sy_dmul(den, q), 1.0);
        const double x = sy_ddiv(num, den);
        return upper ? -x : x;
    }
    const double q = sy_dadd(p, -0.5);
    const double r = sy_dmul(q, q);
    double num = a0;
    num = sy_dadd(sy_dmul(num, r), a1);
    num = sy_dadd(sy_dmul(num, r), a2);
    num = sy_dadd(sy_dmul(num, r), a3);
    num = sy_dadd(sy_dmul(num, r), a4);
    num = sy_dadd(sy_dmul(num, r), a5);
    num = sy_dmul(num, q);
    double den = b0;
    den = sy_dadd(sy_dmul(den, r), b1);
    den = sy_dadd(sy_dmul(den, r), b2);
    den = sy_dadd(sy_dmul(den, r), b3);
    den = sy_dadd(sy_dmul(den, r), b4);
    den = sy_dadd(sy_dmul(den, r), 1.0);
    return sy_ddiv(num, den);
}
SY_HD double sy_normal_d(uint64_t h) { return sy_ndtri(sy_unif_d(h)); }
SY_HD float sy_normal(uint64_t h) { return (float)sy_normal_d(h); }

// ---- negative binomial ------------------------------------------------------------------------------------
SY_HD float sy_p0(float lam)  // P(count = 0) = (r / (r + lam))^r with r = 10
{
    const float q0 = sy_fdiv(SY_SIZE, sy_fadd(SY_SIZE, lam));
    const float q2 = sy_fmul(q0, q0), q4 = sy_fmul(q2, q2), q5 = sy_fmul(q4, q0);
    return sy_fmul(q5, q5);
}
// NB(mean lam, size r) by inversion of u; normal approximation for large means
SY_HD int sy_draw(float lam, float u)
{
    const float r = SY_SIZE;
    if (lam > 30.f) {
        const float sd = sy_fsqrt(sy_fadd(lam, sy_fdiv(sy_fmul(lam, lam), r)));
        const float v = sy_fadd(lam, sy_fmul(sd, (float)sy_ndtri((double)u)));
        return v < 0.f ? 0 : (int)sy_fadd(v, 0.5f);
    }
    float p = sy_p0(lam);
    if (u <= p)
        return 0;
    const float q = sy_fdiv(lam, sy_fadd(r, lam));
    float cum = p;
    int kk = 0;
    while (u > cum && kk < 4096) {
        p = sy_fmul(p, sy_fmul(sy_fdiv(sy_fadd((float)kk, r), sy_fadd((float)kk, 1.f)), q));
        cum = sy_fadd(cum, p);
        ++kk;
        if (p < 1e-12f)
            break;
    }
    return kk;
}

// ---- model parameters (host-computed, G-sized; the device generator receives a copy) -----------------------
SY_HD float sy_gene_base(uint64_t seed, int64_t g) { return sy_fmul(2.0f, sy_normal(sy_hash(seed, 0x1001, (uint64_t)g))); }
SY_HD int sy_gene_factor(uint64_t seed, int64_t g, int z) { return (int)(sy_hash(seed, 0x2000 + z, (uint64_t)g) % SY_F); }
SY_HD float sy_gene_loading(uint64_t seed, int64_t g, int z)
{
    const uint64_t h = sy_hash(seed, 0x2000 + z, (uint64_t)g);
    const int f = (int)(h % SY_F);
    float s = 1.0f;
    for (int i = 0; i < f; ++i)
        s = sy_fmul(s, 0.93f);
    return sy_fmul(sy_normal(sy_mix(h)), s);
}
SY_HD float sy_centroid(uint64_t seed, int i) { return sy_normal(sy_hash(seed, 0x3003, (uint64_t)i)); }

// ---- per-cell latent state ---------------------------------------------------------------------------------
SY_HD int sy_cell_cluster(uint64_t seed, uint64_t cg) { return (int)(sy_hash(seed, 0x4004, cg) % SY_CLUSTERS); }
SY_HD float sy_cell_factor(uint64_t seed, uint64_t cg, int f, const float *cent /* [SY_CLUSTERS*SY_F] */, int cl)
{
    return sy_fadd(cent[cl * SY_F + f], sy_fmul(0.3f, sy_normal(sy_hash(seed, 0x5000 + f, cg))));
}
SY_HD float sy_cell_logdepth(uint64_t seed, uint64_t cg) { return sy_fmul(0.35f, sy_normal(sy_hash(seed, 0x6006, cg))); }

// the count of (gene g, global cell cg); zs = the cell's SY_F latent factors
SY_HD int sy_count(uint64_t seed, uint64_t cg, int64_t g, float base_g, const uint8_t *wf_g, const float *wv_g,
                   const float *zs, float off, float ldepth)
{
    float ll = sy_fadd(sy_fadd(base_g, off), ldepth);
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int z = 0; z < SY_NZ; ++z)
        ll = sy_fadd(ll, sy_fmul(wv_g[z], zs[wf_g[z]]));
    const float lam = sy_expf(ll < SY_LL_MAX ? ll : SY_LL_MAX);
    const float u = sy_unif(sy_hash(seed ^ 0x7007, cg, (uint64_t)g));
    return sy_draw(lam, u);
}

// ---- cfg-4 style embedding: Gaussian clusters, per-PC scale decaying 20 -> 1 ------------------------------
// value(i, t) = scale_t * (centre[cl(i)][t] + 0.35 * N(0,1)), scale_t = 20 / (1 + 19 t / (d - 1))
SY_HD double sy_emb_scale(int t, int d)
{
    return d > 1 ? sy_ddiv(20.0, sy_dadd(1.0, sy_ddiv(sy_dmul(19.0, (double)t), (double)(d - 1)))) : 20.0;
}
SY_HD double sy_emb_value(uint64_t seed, int64_t i, int t, double scale_t, int n_clusters)
{
    const int cl = (int)(sy_hash(seed, 0x8008, (uint64_t)i) % (uint64_t)n_clusters);
    const double cen = sy_normal_d(sy_hash(seed, 0x9000 + (uint64_t)t, (uint64_t)cl));
    const double eps = sy_normal_d(sy_hash(seed, 0xA000 + (uint64_t)t, (uint64_t)i));
    return sy_dmul(scale_t, sy_dadd(cen, sy_dmul(0.35, eps)));
}
