// pshadow_fft2.cuh -- FFT flavour of the filter scan (PSH_MODE_FFT), round 2.  Included by pshadow.cu.
//
// The cross term D_t = sum_j q_j y_{t+j} of ||q - y_t||^2 = Q2 + Y2_t - 2 D_t is a correlation: for a
// whole trajectory it costs O(log N) per window through the FFT instead of W FMAs.  Two trajectories
// share one complex transform (z = y_a + i y_b; q is real, so corr(z, q) = corr(y_a, q) + i corr(y_b, q)).
// What is query-independent is computed once per dataset (psh_fft_prepare):
//   * Zh  : the 4096-point spectra of all row pairs, quantised to fp16 pairs with a per-pair power-of-two
//           scale (4 bytes per complex point);  the quantisation error ||Z^ - Z||_2 / sqrt(N) is MEASURED
//           per pair (zqerr) and enters the rigorous slack through Cauchy-Schwarz:
//           |D^_t - D_t| <= (1/N) sum_k |dZ_k| |Q_k| <= ||dZ||_2 ||q||_2 / sqrt(N);
//   * Y2h : the window energies Y2[r][t], scaled by a per-pair power of two and rounded DOWN to fp16
//           (a lower bound stays a lower bound; the upper bound gives the 2^-10 back), the two rows of a
//           pair interleaved in the scan's output order (4 bytes per pair of windows);
//   * per pair: ||y_pair|| (rounded up), zqerr, the energy scale es and m2 = -2 es / zs.
// Together 8 bytes per pair of samples = exactly the bytes of the raw rows: the scan streams the
// ALGORITHMIC bytes (round 1 streamed fp32 spectra + fp32 energies, 2.05 x).
//
// Per query the scan multiplies by conj(FFT(q))/N, runs ONE inverse 4096-point FFT per row pair in
// registers / shared memory and tests, in the pair's scaled units,
//     fma(m2, v, yf) <= rhs,   rhs = (thr - (Q2 - slack)) * es,
// i.e. the rigorous lower bound LB = Q2 + Y2^ - 2 D^ - slack <= thr of round 1; survivors go through
// the exact re-rank, so results stay bit-identical to the exact scan.
//
// Arithmetic: Blackwell's packed fp32 pipe (add/mul/fma.rn.f32x2 -> SASS FADD2 / FMUL2 / FFMA2).  A
// complex value lives in one 64-bit register pair; a complex add is ONE instruction, a multiplication
// by +-i folds into an FFMA2 with a swapped / half-negated operand, a complex multiplication is two
// (FMUL2 with a broadcast operand + FFMA2): the radix-8 butterfly issues 28 instructions instead of
// 56.  Each lane is an IEEE fp32 operation, so the error analysis of round 1 is unchanged.
//
// FFT: N = 4096 = 8 x 8 x 8 x 8, 512 threads, 8 complex values per thread, four radix-8 passes in
// registers (64 registers per thread: 32 resident warps per SM -- the radix-16 / 256-thread version of
// round 1 ran 16 warps per SM at 128 registers and was latency-bound at 36-64 % issue utilisation).
// Exchange 1 goes through shared memory behind the transform's only CTA barrier, exchange 2 stays inside
// groups of 64 threads (named barriers), exchange 3 is an 8 x 8 transpose inside 8 lanes.  Input index
// n = tid + 512 i; output v[d] = X[kb(tid) + 512 d] (the energies are stored in that order).
#pragma once
// (pshadow.cu includes <cuda_fp16.h> at file scope)

namespace fx2 {

constexpr int N = 4096;
constexpr int THREADS = 512;

#ifndef PSH_SCALAR_FFT
#define PSH_PK2(name, op)                                                                              \
    __device__ __forceinline__ float2 name(float2 a, float2 b) {                                       \
        float2 r;                                                                                      \
        asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; " op                     \
            " rc, ra, rb; mov.b64 {%0,%1}, rc;}"                                                       \
            : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));                          \
        return r;                                                                                      \
    }
PSH_PK2(add2, "add.rn.f32x2")
PSH_PK2(sub2, "sub.rn.f32x2")
PSH_PK2(mul2, "mul.rn.f32x2")
#undef PSH_PK2
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; "
        "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
#else   // A/B build: the same transform on the scalar fp32 pipe (FADD / FMUL / FFMA)
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#endif
// a + i b and a - i b: one FFMA2 each (operand b swapped, one lane negated)
__device__ __forceinline__ float2 add_i(float2 a, float2 b) { return fma2(make_float2(b.y, b.x), make_float2(-1.0f, 1.0f), a); }
__device__ __forceinline__ float2 sub_i(float2 a, float2 b) { return fma2(make_float2(b.y, b.x), make_float2(1.0f, -1.0f), a); }
// a * w = a.x (w.x, w.y) + a.y (-w.y, w.x): FMUL2 (broadcast a.y, swapped w) + FFMA2 (broadcast a.x, addend half-negated)
__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
    const float2 p = mul2(make_float2(a.y, a.y), make_float2(w.y, w.x));
    return fma2(make_float2(a.x, a.x), w, make_float2(-p.x, p.y));
}

// inverse 4-point DFT (e^{+2 pi i nk/4}), in place, natural order
__device__ __forceinline__ void ifft4(float2 &a0, float2 &a1, float2 &a2, float2 &a3) {
    const float2 t0 = add2(a0, a2), t1 = sub2(a0, a2), t2 = add2(a1, a3), t3 = sub2(a1, a3);
    a0 = add2(t0, t2);
    a2 = sub2(t0, t2);
    a1 = add_i(t1, t3);
    a3 = sub_i(t1, t3);
}
// the same with a2 standing for i * a2 (the w16^4 twiddle folded into the butterfly)
__device__ __forceinline__ void ifft4_a2i(float2 &a0, float2 &a1, float2 &a2, float2 &a3) {
    const float2 t0 = add_i(a0, a2), t1 = sub_i(a0, a2), t2 = add2(a1, a3), t3 = sub2(a1, a3);
    a0 = add2(t0, t2);
    a2 = sub2(t0, t2);
    a1 = add_i(t1, t3);
    a3 = sub_i(t1, t3);
}

// inverse 8-point DFT of v[0..7], natural order in and out; 28 packed instructions.
//   X[2m]   = IDFT4(x_j + x_{j+4})_m,   X[2m+1] = IDFT4((x_j - x_{j+4}) w8^j)_m   (w8^2 = i folded into the butterfly)
__device__ __forceinline__ void ifft8(float2 (&v)[8]) {
    const float r = 0.70710678118654752f;
    float2 s0 = add2(v[0], v[4]), d0 = sub2(v[0], v[4]);
    float2 s1 = add2(v[1], v[5]), d1 = sub2(v[1], v[5]);
    float2 s2 = add2(v[2], v[6]), d2 = sub2(v[2], v[6]);
    float2 s3 = add2(v[3], v[7]), d3 = sub2(v[3], v[7]);
    d1 = cmul(d1, make_float2(r, r));      // w8^1
    d3 = cmul(d3, make_float2(-r, r));     // w8^3
    ifft4(s0, s1, s2, s3);                 // X0, X2, X4, X6
    ifft4_a2i(d0, d1, d2, d3);             // X1, X3, X5, X7
    v[0] = s0; v[2] = s1; v[4] = s2; v[6] = s3;
    v[1] = d0; v[3] = d1; v[5] = d2; v[7] = d3;
}

// v[j] *= w^j, j = 1..7, the powers built from the table-exact w^1 (every power is a product of at most
// three values of depth <= 2: error <= ~10 u, inside CF)
__device__ __forceinline__ void twiddle8(float2 (&v)[8], float2 w1) {
    const float2 w2 = cmul(w1, w1), w3 = cmul(w2, w1), w4 = cmul(w2, w2);
    v[1] = cmul(v[1], w1); v[2] = cmul(v[2], w2); v[3] = cmul(v[3], w3); v[4] = cmul(v[4], w4);
    v[5] = cmul(v[5], cmul(w4, w1)); v[6] = cmul(v[6], cmul(w4, w2)); v[7] = cmul(v[7], cmul(w4, w3));
}

// loop-invariant twiddle seeds of a thread: exp(2 pi i tid/4096), exp(2 pi i (tid&63)/512), exp(2 pi i (tid&7)/64)
struct Seeds { float2 s1, s2, s3; };
__device__ __forceinline__ Seeds load_seeds(const float2 *__restrict__ tw, int tid) {
    Seeds s;
    s.s1 = __ldg(tw + tid);
    s.s2 = __ldg(tw + 8 * (tid & 63));
    s.s3 = __ldg(tw + 64 * (tid & 7));
    return s;
}

// Inverse 4096-point transform by one CTA of 512 threads, 8 complex values per thread, four radix-8
// passes (4096 = 8 x 8 x 8 x 8; 64 registers per thread: 32 resident warps per SM).
//   In : v[i] = x[tid + 512 i].
//   Out: v[d] = X[kb + 512 d],  kb = (tid >> 6) + 8 ((tid >> 3) & 7) + 64 (tid & 7)   (tid's octal digits reversed).
// Exchange 1 is CTA-wide (the transform's only CTA barrier; `before_first_barrier()` / `after_first_barrier()`
// run right in front of / behind it -- every thread has consumed the staged spectrum by then); exchange 2 stays
// inside a group of 64 threads (named barrier 1 + group); exchange 3 is an 8 x 8 transpose inside 8 lanes.
// ex1: 8 x 512 float2; ex2: 8 groups x 8 rows x 72 float2 (rows padded: conflict-free 64-bit accesses).
constexpr int EX1_FLOAT2 = 8 * 512;
constexpr int EX2_ROW = 72;
constexpr int EX2_FLOAT2 = 64 * EX2_ROW;

__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <typename F0, typename F>
__device__ __forceinline__ void ifft4096(float2 (&v)[8], float2 *ex1, float2 *ex2, int tid, const Seeds &seeds,
                                         F0 before_first_barrier, F after_first_barrier) {
    ifft8(v);                                   // over n3 -> c
    twiddle8(v, seeds.s1);                      // w4096^(tid c)
#pragma unroll
    for (int c = 0; c < 8; ++c) ex1[c * 512 + tid] = v[c];
    before_first_barrier();
    __syncthreads();
    after_first_barrier();
    const int c = tid >> 6, bp = tid & 63, x = (tid >> 3) & 7, y = tid & 7;
#pragma unroll
    for (int a = 0; a < 8; ++a) v[a] = ex1[c * 512 + 64 * a + bp];
    ifft8(v);                                   // over a' -> c'
    twiddle8(v, seeds.s2);                      // w512^(bp c')
    float2 *g2 = ex2 + c * (8 * EX2_ROW);       // the 64 threads sharing c exchange among themselves
#pragma unroll
    for (int cp = 0; cp < 8; ++cp) g2[cp * EX2_ROW + bp] = v[cp];
    bar_sync_named(1 + c, 64);
    float2 *g3 = g2 + x * EX2_ROW;              // row x is read, then reused, by the 8 lanes (c, x, .) only
#pragma unroll
    for (int a = 0; a < 8; ++a) v[a] = g3[8 * a + y];
    ifft8(v);                                   // over a'' -> c''
    twiddle8(v, seeds.s3);                      // w64^(y c'')
    __syncwarp();
#pragma unroll
    for (int cq = 0; cq < 8; ++cq) g3[cq * 9 + y] = v[cq];
    __syncwarp();
#pragma unroll
    for (int b = 0; b < 8; ++b) v[b] = g3[y * 9 + b];
    ifft8(v);                                   // over b'' -> d
}

// window of output register d of thread tid: kb(tid) + 512 d
__device__ __forceinline__ int out_base(int tid) { return (tid >> 6) + 8 * ((tid >> 3) & 7) + 64 * (tid & 7); }

}  // namespace fx2
