// In-register complex DFTs of small compile-time size (2,3,4,5,8,16,30,32,...), forward sign
// (W_N = exp(-2*pi*i/N)), natural order in and out.  Every index and every twiddle is a compile-time
// constant, so after unrolling the whole transform is straight-line FADD/FMUL/FFMA on registers with
// immediate twiddles — no table, no shared memory, no dynamic register indexing.
//
// The functions are __host__ __device__ so that tests/host_sim can run the very same arithmetic on the CPU
// (there is no GPU in the build container).
#pragma once
#include <cuda_runtime.h>

#include <utility>

#define SELD_HD __host__ __device__ __forceinline__

namespace seld {

// ---- compile-time trigonometry (exact octant reduction, Taylor in double) --------------------------
constexpr double kPi = 3.14159265358979323846264338327950288;

__host__ __device__ constexpr double cx_sin_small(double x) {  // |x| <= pi/4
    double x2 = x * x, term = x, sum = x;
    for (int i = 1; i <= 14; ++i) {
        term *= -x2 / double((2 * i) * (2 * i + 1));
        sum += term;
    }
    return sum;
}
__host__ __device__ constexpr double cx_cos_small(double x) {  // |x| <= pi/4
    double x2 = x * x, term = 1.0, sum = 1.0;
    for (int i = 1; i <= 14; ++i) {
        term *= -x2 / double((2 * i - 1) * (2 * i));
        sum += term;
    }
    return sum;
}
// cos(2*pi*k/n), sin(2*pi*k/n) with the quadrant/octant decided on integers (exact 0, +-1, sqrt(1/2)).
__host__ __device__ constexpr double cx_cos2pi(long long k, long long n) {
    k %= n;
    if (k < 0) k += n;
    if (2 * k > n) k = n - k;                      // cos(2pi - a) = cos a      -> a in [0, pi]
    if (4 * k > n) return -cx_cos2pi(n - 2 * k, 2 * n);  // cos a = -cos(pi - a); pi - a = 2pi (n-2k)/(2n)
    if (8 * k > n) return cx_sin_small(2.0 * kPi * double(n - 4 * k) / double(4 * n));  // cos a = sin(pi/2 - a)
    return cx_cos_small(2.0 * kPi * double(k) / double(n));
}
__host__ __device__ constexpr double cx_sin2pi(long long k, long long n) {
    k %= n;
    if (k < 0) k += n;
    if (2 * k > n) return -cx_sin2pi(n - k, n);          // sin(2pi - a) = -sin a
    if (4 * k > n) return cx_sin2pi(n - 2 * k, 2 * n);   // sin a = sin(pi - a)
    if (8 * k > n) return cx_cos_small(2.0 * kPi * double(n - 4 * k) / double(4 * n));  // sin a = cos(pi/2 - a)
    return cx_sin_small(2.0 * kPi * double(k) / double(n));
}

// ---- static_for ---------------------------------------------------------------------------------
template <class F, int... Is>
SELD_HD void static_for_impl(F&& f, std::integer_sequence<int, Is...>) {
    (f(std::integral_constant<int, Is>{}), ...);
}
template <int N, class F>
SELD_HD void static_for(F&& f) {
    static_for_impl(static_cast<F&&>(f), std::make_integer_sequence<int, N>{});
}

// ---- complex helpers ------------------------------------------------------------------------------
// Complex add / subtract.  On sm_100a these are ONE packed instruction each (add/sub.rn.f32x2 -> FADD2): the FP32
// pipe does the same work, but the butterflies — three quarters of the FFT — take half the issue slots.
// ptxas folds the +-i rotations around them into operand swizzles (.F32x2.LO_HI) and negations.
SELD_HD float2 cadd(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
SELD_HD float2 csub(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; sub.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
#else
    return make_float2(a.x - b.x, a.y - b.y);
#endif
}
// (a.x * w, a.y * w): one FMUL2 with a broadcast scalar operand
SELD_HD float2 cscale(float2 a, float w) {
#ifdef __CUDA_ARCH__
    float2 r;
    asm("{.reg .b64 ra, rw, rc; mov.b64 ra, {%2, %3}; mov.b64 rw, {%4, %4}; mul.rn.f32x2 rc, ra, rw; mov.b64 {%0, %1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(w));
    return r;
#else
    return make_float2(a.x * w, a.y * w);
#endif
}
// Complex multiply a * w.  On sm_100a: TWO packed instructions (FMUL2 + FFMA2) instead of two FMUL + two FFMA — the
// twiddle is one 64-bit operand (.F32x2.HI_LO, and swapped / half-negated .LO_HI.NP for the second product), a.x and a.y
// are scalar-broadcast operands (.F32).  Same roundings as the scalar form.
// (SELD_SCALAR_CMUL, defined by a translation unit before this header, keeps the four-instruction scalar form: the MIC
//  kernel is FP32-pipe-bound rather than issue-bound and measured 7 % slower with the packed form.)
SELD_HD float2 cmul(float2 a, float2 w) {
#if defined(__CUDA_ARCH__) && !defined(SELD_SCALAR_CMUL)
    float2 r;
    asm("{.reg .b64 rw, rws, ra, rb, t; mov.b64 rw, {%4, %5}; mov.b64 rws, {%6, %4}; mov.b64 ra, {%2, %2}; mov.b64 rb, {%3, %3};"
        " mul.rn.f32x2 t, rw, ra; fma.rn.f32x2 t, rws, rb, t; mov.b64 {%0, %1}, t;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(w.x), "f"(w.y), "f"(-w.y));
    return r;
#else
    return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x);
#endif
}
SELD_HD float2 cmul_conj(float2 a, float2 w) { return cmul(a, make_float2(w.x, -w.y)); }
SELD_HD float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)
SELD_HD float2 mul_pi(float2 a) { return make_float2(-a.y, a.x); }  // a * (+i)

// Compile-time twiddles live in constant memory as (c, s, -s, c): ptxas loads them into UNIFORM registers outside the
// frame loop (LDCU) and the packed multiply takes them as 64-bit uniform operands, so a * W_N^K is the same two packed
// instructions with no vector register spent on the constant.
struct TwQuad { float c, s, ns, c2; };
template <int N>
struct TwTable { TwQuad q[N]; };
template <int N>
constexpr TwTable<N> make_tw_table(bool inv) {
    TwTable<N> t{};
    for (int k = 0; k < N; ++k) {
        const float c = float(cx_cos2pi(k, N));
        const float s = float(inv ? cx_sin2pi(k, N) : -cx_sin2pi(k, N));
        t.q[k] = TwQuad{c, s, -s, c};
    }
    return t;
}
#ifdef __CUDACC__
template <int N, bool INV>
__constant__ TwTable<N> kTwConst = make_tw_table<N>(INV);
#endif

// a * W_N^K (forward) or a * conj(W_N^K) (INV), K and N compile-time.
template <int K, int N, bool INV = false>
SELD_HD float2 mul_w(float2 a) {
    constexpr int k = ((K % N) + N) % N;
    if constexpr (k == 0) {
        return a;
    } else if constexpr (4 * k == N) {
        return INV ? mul_pi(a) : mul_mi(a);
    } else if constexpr (2 * k == N) {
        return make_float2(-a.x, -a.y);
    } else if constexpr (4 * k == 3 * N) {
        return INV ? mul_mi(a) : mul_pi(a);
    } else {
#if defined(__CUDA_ARCH__) && !defined(SELD_SCALAR_CMUL)
        const TwQuad& q = kTwConst<N, INV>.q[k];
        float2 r;
        asm("{.reg .b64 rw, rws, ra, rb, t; mov.b64 rw, {%4, %5}; mov.b64 rws, {%6, %7}; mov.b64 ra, {%2, %2}; mov.b64 rb, {%3, %3};"
            " mul.rn.f32x2 t, rw, ra; fma.rn.f32x2 t, rws, rb, t; mov.b64 {%0, %1}, t;}"
            : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(q.c), "f"(q.s), "f"(q.ns), "f"(q.c2));
        return r;
#else
        constexpr float c = float(cx_cos2pi(k, N));
        constexpr float s = float(INV ? cx_sin2pi(k, N) : -cx_sin2pi(k, N));
        return make_float2(a.x * c - a.y * s, a.x * s + a.y * c);
#endif
    }
}

// ---- prime-size kernels ------------------------------------------------------------------------
template <int N, bool INV = false>
struct Dft;

template <bool INV>
struct Dft<1, INV> {
    static SELD_HD void run(float2 (&)[1]) {}
};

template <bool INV>
struct Dft<2, INV> {
    static SELD_HD void run(float2 (&v)[2]) {
        float2 a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};

template <bool INV>
struct Dft<3, INV> {
    static SELD_HD void run(float2 (&v)[3]) {
        constexpr float c = float(cx_cos2pi(1, 3));                               // -0.5
        constexpr float s = float(INV ? cx_sin2pi(1, 3) : -cx_sin2pi(1, 3));      // -+sqrt(3)/2
        float2 t1 = cadd(v[1], v[2]);
        float2 t2 = csub(v[1], v[2]);
        float2 m = make_float2(v[0].x + c * t1.x, v[0].y + c * t1.y);
        float2 r = make_float2(-s * t2.y, s * t2.x);  // i*s*t2
        v[0] = cadd(v[0], t1);
        v[1] = cadd(m, r);
        v[2] = csub(m, r);
    }
};

template <bool INV>
struct Dft<4, INV> {
    static SELD_HD void run(float2 (&v)[4]) {
        float2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
        float2 t2 = cadd(v[1], v[3]), t3 = csub(v[1], v[3]);
        float2 r = INV ? mul_pi(t3) : mul_mi(t3);
        v[0] = cadd(t0, t2);
        v[2] = csub(t0, t2);
        v[1] = cadd(t1, r);
        v[3] = csub(t1, r);
    }
};

template <bool INV>
struct Dft<5, INV> {
    static SELD_HD void run(float2 (&v)[5]) {
        constexpr float c1 = float(cx_cos2pi(1, 5)), c2 = float(cx_cos2pi(2, 5));
        constexpr float s1 = float(INV ? cx_sin2pi(1, 5) : -cx_sin2pi(1, 5));
        constexpr float s2 = float(INV ? cx_sin2pi(2, 5) : -cx_sin2pi(2, 5));
        float2 a1 = cadd(v[1], v[4]), b1 = csub(v[1], v[4]);
        float2 a2 = cadd(v[2], v[3]), b2 = csub(v[2], v[3]);
        float2 m1 = make_float2(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
        float2 m2 = make_float2(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
        // i*(s1*b1 + s2*b2) and i*(s2*b1 - s1*b2)
        float2 r1 = make_float2(-(s1 * b1.y + s2 * b2.y), s1 * b1.x + s2 * b2.x);
        float2 r2 = make_float2(-(s2 * b1.y - s1 * b2.y), s2 * b1.x - s1 * b2.x);
        v[0] = make_float2(v[0].x + a1.x + a2.x, v[0].y + a1.y + a2.y);
        v[1] = cadd(m1, r1);
        v[4] = csub(m1, r1);
        v[2] = cadd(m2, r2);
        v[3] = csub(m2, r2);
    }
};

// ---- Cooley-Tukey composite: N = A * B --------------------------------------------------------
// n = j + B*a (j<B, a<A), k = q + A*b (q<A, b<B):
//   u[j][q] = sum_a x[j+B*a] W_A^{aq};  u *= W_N^{jq};  X[q+A*b] = sum_j u[j][q] W_B^{jb}.
template <int A, int B, bool INV>
struct DftCT {
    static constexpr int N = A * B;
    static SELD_HD void run(float2 (&v)[N]) {
        float2 w[N];  // w[q*B + j]
        static_for<B>([&](auto J) {
            constexpr int j = decltype(J)::value;
            float2 t[A];
            static_for<A>([&](auto Ai) { constexpr int a = decltype(Ai)::value; t[a] = v[j + B * a]; });
            Dft<A, INV>::run(t);
            static_for<A>([&](auto Q) {
                constexpr int q = decltype(Q)::value;
                w[q * B + j] = mul_w<j * q, N, INV>(t[q]);
            });
        });
        static_for<A>([&](auto Q) {
            constexpr int q = decltype(Q)::value;
            float2 t[B];
            static_for<B>([&](auto J) { constexpr int j = decltype(J)::value; t[j] = w[q * B + j]; });
            Dft<B, INV>::run(t);
            static_for<B>([&](auto Bi) { constexpr int b = decltype(Bi)::value; v[q + A * b] = t[b]; });
        });
    }
};

template <bool INV> struct Dft<6, INV>  { static SELD_HD void run(float2 (&v)[6])  { DftCT<2, 3, INV>::run(v); } };
template <bool INV> struct Dft<8, INV>  { static SELD_HD void run(float2 (&v)[8])  { DftCT<2, 4, INV>::run(v); } };
template <bool INV> struct Dft<10, INV> { static SELD_HD void run(float2 (&v)[10]) { DftCT<2, 5, INV>::run(v); } };
template <bool INV> struct Dft<15, INV> { static SELD_HD void run(float2 (&v)[15]) { DftCT<3, 5, INV>::run(v); } };
template <bool INV> struct Dft<16, INV> { static SELD_HD void run(float2 (&v)[16]) { DftCT<4, 4, INV>::run(v); } };
template <bool INV> struct Dft<30, INV> { static SELD_HD void run(float2 (&v)[30]) { DftCT<5, 6, INV>::run(v); } };
template <bool INV> struct Dft<32, INV> { static SELD_HD void run(float2 (&v)[32]) { DftCT<4, 8, INV>::run(v); } };

}  // namespace seld
