// glsl_pk.h -- "two rays per lane" value types for the march kernels (device only, sm_100a).
//
// Included INSIDE the scene namespace (xg or fg) after glsl_rt.h, by raymarch_kernel.cuh when RM_DUAL
// is on.  A `pf` holds the same quantity for TWO rays (lo, hi) in one 64-bit register pair; pvec2/3/4
// are vectors of pf.  Arithmetic issues as Blackwell packed FP32 (FADD2 / FMUL2 / FFMA2: two IEEE
// round-to-nearest operations per lane per issue slot), everything that has no packed form (floor,
// min/max, abs, square-root seed, transcendentals) runs per half with the scalar built-ins of the
// enclosing namespace.  Every function performs, per half, exactly the operation sequence of its
// scalar twin in glsl_rt.h, so the exact flavour stays bit-identical to the one-ray code (and to
// the CPU oracle).
//
// Scene code reaches these types through the "varying" lowering (lower_glsl.cpp, lower_scene_packed):
// functions become templates over their parameter types, initialised locals become `auto`, so uniform-
// only sub-expressions stay plain `float` (and fold at compile time) while everything derived from the
// ray position becomes packed.  Constructs with no packed meaning (comparisons and branches on varying
// values, swizzles, matrices) do not compile; the host then falls back to the one-ray kernels.
#ifndef RM_PK_FAST
#error "define RM_PK_FAST (0 exact / 1 fast) before including glsl_pk.h"
#endif

extern "C" {
__device__ __device_builtin__ float2 __ffma2_rn_impl(float2 x, float2 y, float2 z);
__device__ __device_builtin__ float2 __fadd2_rn_impl(float2 x, float2 y);
__device__ __device_builtin__ float2 __fmul2_rn_impl(float2 x, float2 y);
__device__ __device_builtin__ float2 __fadd2_rd_impl(float2 x, float2 y);
}

struct pf {
    float2 v;
    RM_HD pf() {}
    RM_HD pf(float a) { v = make_float2(a, a); }                 // uniform value, same for both rays
    RM_HD pf(float a, float b) { v = make_float2(a, b); }
    RM_HD float lo() const { return v.x; }
    RM_HD float hi() const { return v.y; }
};
RM_HD pf pk(float2 v) { pf r; r.v = v; return r; }
RM_HD float2 pk_neg(float2 a) { return make_float2(-a.x, -a.y); }

// ---- arithmetic (packed) ---------------------------------------------------------------------
RM_HD pf operator+(const pf& a, const pf& b) { return pk(__fadd2_rn_impl(a.v, b.v)); }
RM_HD pf operator-(const pf& a, const pf& b) { return pk(__fadd2_rn_impl(a.v, pk_neg(b.v))); }   // a + (-b) == a - b
RM_HD pf operator*(const pf& a, const pf& b) { return pk(__fmul2_rn_impl(a.v, b.v)); }
RM_HD pf operator-(const pf& a) { return pk(pk_neg(a.v)); }
RM_HD pf operator+(const pf& a) { return a; }
RM_HD pf g_fma(const pf& a, const pf& b, const pf& c) { return pk(__ffma2_rn_impl(a.v, b.v, c.v)); }
// per-half helpers for everything else
#define RM_PK1(expr_lo, expr_hi) pf((expr_lo), (expr_hi))
RM_HD pf g_rcp(const pf& a) { return RM_PK1(g_rcp(a.v.x), g_rcp(a.v.y)); }
RM_HD pf operator/(const pf& a, const pf& b) { return RM_PK1(g_div(a.v.x, b.v.x), g_div(a.v.y, b.v.y)); }   // scalar / scalar: IEEE (glsl_rt.h)
RM_HD pf& operator+=(pf& a, const pf& b) { a = a + b; return a; }
RM_HD pf& operator-=(pf& a, const pf& b) { a = a - b; return a; }
RM_HD pf& operator*=(pf& a, const pf& b) { a = a * b; return a; }
RM_HD pf& operator/=(pf& a, const pf& b) { a = a / b; return a; }

#define RM_PK_MAP1(name) RM_HD pf name(const pf& a) { return RM_PK1(name(a.v.x), name(a.v.y)); }
#define RM_PK_MAP2(name) RM_HD pf name(const pf& a, const pf& b) { return RM_PK1(name(a.v.x, b.v.x), name(a.v.y, b.v.y)); }
RM_PK_MAP1(radians) RM_PK_MAP1(degrees) RM_PK_MAP1(sin) RM_PK_MAP1(cos) RM_PK_MAP1(tan) RM_PK_MAP1(asin) RM_PK_MAP1(acos) RM_PK_MAP1(atan)
RM_PK_MAP1(sinh) RM_PK_MAP1(cosh) RM_PK_MAP1(tanh) RM_PK_MAP1(asinh) RM_PK_MAP1(acosh) RM_PK_MAP1(atanh)
RM_PK_MAP1(exp) RM_PK_MAP1(log) RM_PK_MAP1(exp2) RM_PK_MAP1(log2) RM_PK_MAP1(sqrt) RM_PK_MAP1(inversesqrt)
RM_PK_MAP1(abs) RM_PK_MAP1(sign) RM_PK_MAP1(floor) RM_PK_MAP1(trunc) RM_PK_MAP1(round) RM_PK_MAP1(roundEven) RM_PK_MAP1(ceil)
RM_PK_MAP2(atan) RM_PK_MAP2(pow) RM_PK_MAP2(min) RM_PK_MAP2(max) RM_PK_MAP2(step)
RM_HD pf fract(const pf& x) { return x - floor(x); }
RM_HD pf mod(const pf& x, const pf& y) { return g_fma(-y, floor(x * g_rcp(y)), x); }
RM_HD pf mod(const pf& x, float y) { return g_fma(pf(-y), floor(x * pf(g_rcp(y))), x); }
RM_HD pf clamp(const pf& x, const pf& lo, const pf& hi) { return min(max(x, lo), hi); }
RM_HD pf mix(const pf& a, const pf& b, const pf& t) { return g_fma(b, t, a * (pf(1.0f) - t)); }
RM_HD pf smoothstep(const pf& e0, const pf& e1, const pf& x) {
    const pf t = clamp((x - e0) / (e1 - e0), pf(0.0f), pf(1.0f));
    return (t * t) * (pf(3.0f) - pf(2.0f) * t);
}
RM_HD pf length(const pf& x) { return abs(x); }
RM_HD pf distance(const pf& a, const pf& b) { return abs(a - b); }
RM_HD pf dot(const pf& a, const pf& b) { return a * b; }
// domain repetition of one component (rm_rep1 / rm_rep0 of glsl_rt.h)
#if RM_PK_FAST
RM_HD pf rm_rep1(const pf& x, float h1, float s, float h2) {
    if (h2 == 0.5f * s && s > 0.0f && s < 1e30f) {
        const float M = 12582912.0f;
        const pf y = (h1 == h2) ? x : x + pf(h1 - h2);
        const pf r = g_fma(y, pf(1.0f / s), pf(M)) + pf(-M);
        return g_fma(pf(-s), r, y);
    }
    return mod(x + pf(h1), s) - pf(h2);
}
RM_HD pf rm_rep0(const pf& x, float s, float h2) { return rm_rep1(x, 0.0f, s, h2); }
#else
RM_HD pf rm_rep1(const pf& x, float h1, float s, float h2) { return mod(x + pf(h1), s) - pf(h2); }
RM_HD pf rm_rep0(const pf& x, float s, float h2) { return mod(x, s) - pf(h2); }
#endif
RM_HD pf rm_rep(const pf& x, float h1, float s, float h2) { return rm_rep1(x, h1, s, h2); }

// ---- vectors of pf -----------------------------------------------------------------------------
struct pvec2 {
    pf x, y;
    RM_HD pvec2() {}
    RM_HD pvec2(const pf& a, const pf& b) : x(a), y(b) {}
    RM_HD explicit pvec2(const pf& a) : x(a), y(a) {}
    RM_HD pvec2(const vec2& u) : x(u.x), y(u.y) {}
};
struct pvec3 {
    pf x, y, z;
    RM_HD pvec3() {}
    RM_HD pvec3(const pf& a, const pf& b, const pf& c) : x(a), y(b), z(c) {}
    RM_HD explicit pvec3(const pf& a) : x(a), y(a), z(a) {}
    RM_HD pvec3(const vec3& u) : x(u.x), y(u.y), z(u.z) {}
    RM_HD pvec3(const pvec2& a, const pf& c) : x(a.x), y(a.y), z(c) {}
};
struct pvec4 {
    pf x, y, z, w;
    RM_HD pvec4() {}
    RM_HD pvec4(const pf& a, const pf& b, const pf& c, const pf& d) : x(a), y(b), z(c), w(d) {}
    RM_HD explicit pvec4(const pf& a) : x(a), y(a), z(a), w(a) {}
    RM_HD pvec4(const vec4& u) : x(u.x), y(u.y), z(u.z), w(u.w) {}
    RM_HD pvec4(const pvec3& a, const pf& d) : x(a.x), y(a.y), z(a.z), w(d) {}
};
// GLSL constructors spelled with the GLSL names resolve here when an argument is packed
RM_HD pvec2 vec2_(const pf& a, const pf& b) { return pvec2(a, b); }
RM_HD pvec3 vec3_(const pf& a, const pf& b, const pf& c) { return pvec3(a, b, c); }

#define RM_PK_VBIN(op)                                                                                              \
    RM_HD pvec2 operator op(const pvec2& a, const pvec2& b) { return pvec2(a.x op b.x, a.y op b.y); }               \
    RM_HD pvec2 operator op(const pvec2& a, const pf& b) { return pvec2(a.x op b, a.y op b); }                      \
    RM_HD pvec2 operator op(const pf& a, const pvec2& b) { return pvec2(a op b.x, a op b.y); }                      \
    RM_HD pvec3 operator op(const pvec3& a, const pvec3& b) { return pvec3(a.x op b.x, a.y op b.y, a.z op b.z); }   \
    RM_HD pvec3 operator op(const pvec3& a, const pf& b) { return pvec3(a.x op b, a.y op b, a.z op b); }            \
    RM_HD pvec3 operator op(const pf& a, const pvec3& b) { return pvec3(a op b.x, a op b.y, a op b.z); }            \
    RM_HD pvec4 operator op(const pvec4& a, const pvec4& b) { return pvec4(a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w); } \
    RM_HD pvec4 operator op(const pvec4& a, const pf& b) { return pvec4(a.x op b, a.y op b, a.z op b, a.w op b); }  \
    RM_HD pvec4 operator op(const pf& a, const pvec4& b) { return pvec4(a op b.x, a op b.y, a op b.z, a op b.w); }  \
    RM_HD pvec2& operator op##=(pvec2& a, const pvec2& b) { a = a op b; return a; }                                 \
    RM_HD pvec2& operator op##=(pvec2& a, const pf& b) { a = a op b; return a; }                                    \
    RM_HD pvec3& operator op##=(pvec3& a, const pvec3& b) { a = a op b; return a; }                                 \
    RM_HD pvec3& operator op##=(pvec3& a, const pf& b) { a = a op b; return a; }                                    \
    RM_HD pvec4& operator op##=(pvec4& a, const pvec4& b) { a = a op b; return a; }                                 \
    RM_HD pvec4& operator op##=(pvec4& a, const pf& b) { a = a op b; return a; }
RM_PK_VBIN(+)
RM_PK_VBIN(-)
RM_PK_VBIN(*)
#undef RM_PK_VBIN
// vector / vector: IEEE per component; vector / scalar: times the correctly rounded reciprocal (glsl_rt.h)
RM_HD pvec2 operator/(const pvec2& a, const pvec2& b) { return pvec2(a.x / b.x, a.y / b.y); }
RM_HD pvec3 operator/(const pvec3& a, const pvec3& b) { return pvec3(a.x / b.x, a.y / b.y, a.z / b.z); }
RM_HD pvec4 operator/(const pvec4& a, const pvec4& b) { return pvec4(a.x / b.x, a.y / b.y, a.z / b.z, a.w / b.w); }
RM_HD pvec2 operator/(const pvec2& a, const pf& b) { const pf r = g_rcp(b); return pvec2(a.x * r, a.y * r); }
RM_HD pvec3 operator/(const pvec3& a, const pf& b) { const pf r = g_rcp(b); return pvec3(a.x * r, a.y * r, a.z * r); }
RM_HD pvec4 operator/(const pvec4& a, const pf& b) { const pf r = g_rcp(b); return pvec4(a.x * r, a.y * r, a.z * r, a.w * r); }
RM_HD pvec2 operator/(const pf& a, const pvec2& b) { return pvec2(a / b.x, a / b.y); }
RM_HD pvec3 operator/(const pf& a, const pvec3& b) { return pvec3(a / b.x, a / b.y, a / b.z); }
RM_HD pvec3& operator/=(pvec3& a, const pf& b) { a = a / b; return a; }
RM_HD pvec3& operator/=(pvec3& a, const pvec3& b) { a = a / b; return a; }
RM_HD pvec2 operator-(const pvec2& a) { return pvec2(-a.x, -a.y); }
RM_HD pvec3 operator-(const pvec3& a) { return pvec3(-a.x, -a.y, -a.z); }
RM_HD pvec4 operator-(const pvec4& a) { return pvec4(-a.x, -a.y, -a.z, -a.w); }

#define RM_PK_VMAP1(name)                                                                    \
    RM_HD pvec2 name(const pvec2& a) { return pvec2(name(a.x), name(a.y)); }                 \
    RM_HD pvec3 name(const pvec3& a) { return pvec3(name(a.x), name(a.y), name(a.z)); }      \
    RM_HD pvec4 name(const pvec4& a) { return pvec4(name(a.x), name(a.y), name(a.z), name(a.w)); }
#define RM_PK_VMAP2(name)                                                                                          \
    RM_HD pvec2 name(const pvec2& a, const pvec2& b) { return pvec2(name(a.x, b.x), name(a.y, b.y)); }             \
    RM_HD pvec3 name(const pvec3& a, const pvec3& b) { return pvec3(name(a.x, b.x), name(a.y, b.y), name(a.z, b.z)); } \
    RM_HD pvec4 name(const pvec4& a, const pvec4& b) { return pvec4(name(a.x, b.x), name(a.y, b.y), name(a.z, b.z), name(a.w, b.w)); } \
    RM_HD pvec2 name(const pvec2& a, const pf& b) { return pvec2(name(a.x, b), name(a.y, b)); }                    \
    RM_HD pvec3 name(const pvec3& a, const pf& b) { return pvec3(name(a.x, b), name(a.y, b), name(a.z, b)); }      \
    RM_HD pvec4 name(const pvec4& a, const pf& b) { return pvec4(name(a.x, b), name(a.y, b), name(a.z, b), name(a.w, b)); }
RM_PK_VMAP1(radians) RM_PK_VMAP1(degrees) RM_PK_VMAP1(sin) RM_PK_VMAP1(cos) RM_PK_VMAP1(tan) RM_PK_VMAP1(asin) RM_PK_VMAP1(acos) RM_PK_VMAP1(atan)
RM_PK_VMAP1(exp) RM_PK_VMAP1(log) RM_PK_VMAP1(exp2) RM_PK_VMAP1(log2) RM_PK_VMAP1(sqrt) RM_PK_VMAP1(inversesqrt)
RM_PK_VMAP1(sign) RM_PK_VMAP1(floor) RM_PK_VMAP1(trunc) RM_PK_VMAP1(round) RM_PK_VMAP1(roundEven) RM_PK_VMAP1(ceil) RM_PK_VMAP1(fract)
RM_PK_VMAP2(pow) RM_PK_VMAP2(min) RM_PK_VMAP2(max) RM_PK_VMAP2(mod)
#undef RM_PK_VMAP1
#undef RM_PK_VMAP2
RM_HD pvec3 mod(const pvec3& a, float b) { return pvec3(mod(a.x, b), mod(a.y, b), mod(a.z, b)); }
RM_HD pvec2 mod(const pvec2& a, float b) { return pvec2(mod(a.x, b), mod(a.y, b)); }
RM_HD pvec3 clamp(const pvec3& v, const pf& lo, const pf& hi) { return pvec3(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi)); }
RM_HD pvec3 mix(const pvec3& a, const pvec3& b, const pf& t) { return pvec3(mix(a.x, b.x, t), mix(a.y, b.y, t), mix(a.z, b.z, t)); }

// abs(v) as a proxy: `abs(v) - b` - the shape every box / repetition SDF has - then runs per half as a
// scalar subtract whose |.| is a free operand modifier, instead of explicit abs instructions feeding a
// packed subtract.  Anything else converts the proxy to the plain packed vector.
struct pabs2 { pvec2 a; RM_HD operator pvec2() const { return pvec2(abs(a.x), abs(a.y)); } };
struct pabs3 { pvec3 a; RM_HD operator pvec3() const { return pvec3(abs(a.x), abs(a.y), abs(a.z)); } };
RM_HD pabs2 abs(const pvec2& a) { pabs2 r; r.a = a; return r; }
RM_HD pabs3 abs(const pvec3& a) { pabs3 r; r.a = a; return r; }
RM_HD pvec4 abs(const pvec4& a) { return pvec4(abs(a.x), abs(a.y), abs(a.z), abs(a.w)); }
RM_HD pf rm_abs_sub(const pf& a, const pf& b) { return pf(g_sub(g_abs(a.v.x), b.v.x), g_sub(g_abs(a.v.y), b.v.y)); }
RM_HD pvec3 operator-(const pabs3& a, const pvec3& b) { return pvec3(rm_abs_sub(a.a.x, b.x), rm_abs_sub(a.a.y, b.y), rm_abs_sub(a.a.z, b.z)); }
RM_HD pvec3 operator-(const pabs3& a, const pf& b) { return pvec3(rm_abs_sub(a.a.x, b), rm_abs_sub(a.a.y, b), rm_abs_sub(a.a.z, b)); }
RM_HD pvec2 operator-(const pabs2& a, const pvec2& b) { return pvec2(rm_abs_sub(a.a.x, b.x), rm_abs_sub(a.a.y, b.y)); }
RM_HD pvec2 operator-(const pabs2& a, const pf& b) { return pvec2(rm_abs_sub(a.a.x, b), rm_abs_sub(a.a.y, b)); }

// ---- geometric ---------------------------------------------------------------------------------
RM_HD pf dot(const pvec2& a, const pvec2& b) { return g_fma(a.y, b.y, a.x * b.x); }
RM_HD pf dot(const pvec3& a, const pvec3& b) { return g_fma(a.z, b.z, g_fma(a.y, b.y, a.x * b.x)); }
RM_HD pf dot(const pvec4& a, const pvec4& b) { return g_fma(a.w, b.w, g_fma(a.z, b.z, g_fma(a.y, b.y, a.x * b.x))); }
RM_HD pf length(const pvec2& a) { return sqrt(dot(a, a)); }
RM_HD pf length(const pvec3& a) { return sqrt(dot(a, a)); }
RM_HD pf length(const pvec4& a) { return sqrt(dot(a, a)); }
RM_HD pf distance(const pvec2& a, const pvec2& b) { return length(a - b); }
RM_HD pf distance(const pvec3& a, const pvec3& b) { return length(a - b); }
#if RM_PK_FAST
RM_HD pvec2 normalize(const pvec2& a) { return a * inversesqrt(dot(a, a)); }
RM_HD pvec3 normalize(const pvec3& a) { return a * inversesqrt(dot(a, a)); }
#else
RM_HD pvec2 normalize(const pvec2& a) { return a / length(a); }
RM_HD pvec3 normalize(const pvec3& a) { return a / length(a); }
#endif
RM_HD pvec3 cross(const pvec3& a, const pvec3& b) {
    return pvec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
RM_HD pvec3 reflect(const pvec3& i, const pvec3& n) { return i - (pf(2.0f) * dot(n, i)) * n; }

// domain repetition of vectors (rm_rep / rm_rep0): operands may be float, pf, vecN or pvecN
RM_HD float rm_c(const pf&, int);   // (never called: packed scales / offsets are not supported, keeps overload sets well-formed)
template <class H1, class S, class H2> RM_HD pvec3 rm_rep(const pvec3& x, const H1& h1, const S& s, const H2& h2) {
    return pvec3(rm_rep1(x.x, rm_c(h1, 0), rm_c(s, 0), rm_c(h2, 0)), rm_rep1(x.y, rm_c(h1, 1), rm_c(s, 1), rm_c(h2, 1)),
                 rm_rep1(x.z, rm_c(h1, 2), rm_c(s, 2), rm_c(h2, 2)));
}
template <class H1, class S, class H2> RM_HD pvec2 rm_rep(const pvec2& x, const H1& h1, const S& s, const H2& h2) {
    return pvec2(rm_rep1(x.x, rm_c(h1, 0), rm_c(s, 0), rm_c(h2, 0)), rm_rep1(x.y, rm_c(h1, 1), rm_c(s, 1), rm_c(h2, 1)));
}
template <class S, class H2> RM_HD pvec3 rm_rep0(const pvec3& x, const S& s, const H2& h2) {
    return pvec3(rm_rep0(x.x, rm_c(s, 0), rm_c(h2, 0)), rm_rep0(x.y, rm_c(s, 1), rm_c(h2, 1)), rm_rep0(x.z, rm_c(s, 2), rm_c(h2, 2)));
}
template <class S, class H2> RM_HD pvec2 rm_rep0(const pvec2& x, const S& s, const H2& h2) {
    return pvec2(rm_rep0(x.x, rm_c(s, 0), rm_c(h2, 0)), rm_rep0(x.y, rm_c(s, 1), rm_c(h2, 1)));
}
#undef RM_PK_MAP1
#undef RM_PK_MAP2
