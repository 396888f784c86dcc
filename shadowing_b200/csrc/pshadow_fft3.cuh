// pshadow_fft3.cuh -- FFT flavour of the filter scan on WARP-AUTONOMOUS 1024-point transforms.
// Included by pshadow.cu behind pshadow_fftscan.cuh (formats, thresholds and the rigorous bound are the
// ones documented in pshadow_fft2.cuh / pshadow_fftscan.cuh; only the transform size and who runs it differ).
//
// Why.  The 4096-point kernel (fft_scan_kernel) runs one transform per CTA of 8 warps: its all-to-all
// exchange needs a CTA barrier, the two resident CTAs of an SM sit in the same phase most of the time, and
// ncu shows the result -- the packed fp32 pipe saturated inside the radix-16 passes and idle in between
// (47 % busy overall, 25 % of all warp samples waiting at the barriers).  Here ONE WARP owns a transform:
//   * N = 1024 = 32 x 32: two radix-32 passes in registers (32 complex values per lane) and ONE 32 x 32
//     transpose through a warp-private shared-memory tile (__syncwarp only) -- no CTA barrier anywhere in
//     the loop, 16 warps per SM all in different phases, so one warp's loads, transposes and epilogue
//     overlap the others' butterflies;
//   * inter-pass twiddles come from an exact table in shared memory (one LDS.128 per two twiddles) instead
//     of products of seeds: per point the packed-pipe work drops from 25 to 19 instructions;
//   * a trajectory is cut into overlapping 1024-sample pieces (overlap-save; "virtual rows", as the
//     4096-point flavour does for T > 4096): a piece yields 1024 - W + 1 windows, so the spectra are
//     1024 / (1025 - W) times the raw rows (1.33 x at W = 252) while the energies stay 1.0 x.
// Layouts (psh_fft_prepare): the spectrum of a pair and its window energies are stored so that a lane
// fetches its 32 values with eight conflict-free 16-byte loads: element e = lane + 32 i (frequency or
// window) of a 4-byte array lives at ((i >> 2) * 32 + lane) * 4 + (i & 3); the query spectrum (8-byte
// elements) at ((i >> 1) * 32 + lane) * 2 + (i & 1).
#pragma once

// exp(2 pi i n / 32): the butterflies' constant twiddles as uniform operands of the packed instructions (constant bank ->
// uniform register) instead of immediates moved into vector registers on the fp32 pipe
__constant__ float2 c_w32[32] = {
    {1.000000000e+00f, 0.000000000e+00f},
    {9.807852804e-01f, 1.950903220e-01f},
    {9.238795325e-01f, 3.826834324e-01f},
    {8.314696123e-01f, 5.555702330e-01f},
    {7.071067812e-01f, 7.071067812e-01f},
    {5.555702330e-01f, 8.314696123e-01f},
    {3.826834324e-01f, 9.238795325e-01f},
    {1.950903220e-01f, 9.807852804e-01f},
    {0.000000000e+00f, 1.000000000e+00f},
    {-1.950903220e-01f, 9.807852804e-01f},
    {-3.826834324e-01f, 9.238795325e-01f},
    {-5.555702330e-01f, 8.314696123e-01f},
    {-7.071067812e-01f, 7.071067812e-01f},
    {-8.314696123e-01f, 5.555702330e-01f},
    {-9.238795325e-01f, 3.826834324e-01f},
    {-9.807852804e-01f, 1.950903220e-01f},
    {-1.000000000e+00f, 1.224646799e-16f},
    {-9.807852804e-01f, -1.950903220e-01f},
    {-9.238795325e-01f, -3.826834324e-01f},
    {-8.314696123e-01f, -5.555702330e-01f},
    {-7.071067812e-01f, -7.071067812e-01f},
    {-5.555702330e-01f, -8.314696123e-01f},
    {-3.826834324e-01f, -9.238795325e-01f},
    {-1.950903220e-01f, -9.807852804e-01f},
    {-1.836970199e-16f, -1.000000000e+00f},
    {1.950903220e-01f, -9.807852804e-01f},
    {3.826834324e-01f, -9.238795325e-01f},
    {5.555702330e-01f, -8.314696123e-01f},
    {7.071067812e-01f, -7.071067812e-01f},
    {8.314696123e-01f, -5.555702330e-01f},
    {9.238795325e-01f, -3.826834324e-01f},
    {9.807852804e-01f, -1.950903220e-01f}};

#ifndef PSH_TWPF
#define PSH_TWPF 3    // twiddle loads in flight
#endif
#ifndef PSH_TWPF0
#define PSH_TWPF0 0   // of which issued before the first pass (measured: 0 / 3 best, 0.1687 -> 0.1663 ms)
#endif

namespace fx3 {

constexpr int N = 1024;
constexpr int EXS = 34;                        // float2 per row of the transpose tile: rows 16-byte aligned, LDS.128 conflict-free
constexpr int EX_BYTES = 32 * EXS * 8;         // 8704
constexpr int Z_BYTES = N * 4;                 // staged spectrum (half2)
constexpr int Y_BYTES = N * 4;                 // staged energies (half2), first ncy * 512 bytes used
constexpr int WARP_BYTES_ALIAS = EX_BYTES + Y_BYTES + 32;            // spectrum staged inside the transpose tile
constexpr int WARP_BYTES_SEP = EX_BYTES + Z_BYTES + Y_BYTES + 32;    // separate spectrum buffer (groups of queries)
constexpr int TW_BYTES = 16 * 32 * 16;         // twiddle table, float4 {w^(lane 2j), w^(lane (2j+1))}
constexpr int Q_BYTES = N * 8;                 // one query's spectrum
#ifndef PSH_FFT3_WARPS
#define PSH_FFT3_WARPS 16
#endif
constexpr int WARPS_SINGLE = PSH_FFT3_WARPS;
constexpr int WARPS_GROUP = 13;            // (13 x 16928 + 8192 bytes: what fits the 227 KB of a CTA)

using fx2::add2;
using fx2::sub2;
using fx2::mul2;
using fx2::fma2;
using fx2::cmul;
using fx2::ifft4;
using fx2::ifft4_a2i;

__host__ __device__ __forceinline__ int perm4(int e) { const int l = e & 31, i = e >> 5; return ((i >> 2) * 32 + l) * 4 + (i & 3); }
__host__ __device__ __forceinline__ int perm2(int e) { const int l = e & 31, i = e >> 5; return ((i >> 1) * 32 + l) * 2 + (i & 1); }
// window / frequency held at position pos of a permuted 4-byte array
__host__ __device__ __forceinline__ int unperm4(int pos) { const int j = pos & 3, l = (pos >> 2) & 31, g = pos >> 7; return l + 32 * (4 * g + j); }

// inverse 16-point DFT, natural order in and out; V8I: v[8] stands for i * v[8]
template <bool V8I>
__device__ __forceinline__ void ifft16(float2 (&v)[16]) {
    if (V8I) ifft4_a2i(v[0], v[4], v[8], v[12]);
    else ifft4(v[0], v[4], v[8], v[12]);
#pragma unroll
    for (int n1 = 1; n1 < 4; ++n1) ifft4(v[n1], v[n1 + 4], v[n1 + 8], v[n1 + 12]);
    v[1 + 4 * 1] = cmul(v[1 + 4 * 1], c_w32[2]);      // w16^1
    v[2 + 4 * 1] = cmul(v[2 + 4 * 1], c_w32[4]);      // w16^2
    v[3 + 4 * 1] = cmul(v[3 + 4 * 1], c_w32[6]);      // w16^3
    v[1 + 4 * 2] = cmul(v[1 + 4 * 2], c_w32[4]);      // w16^2
    //  v[2 + 4 * 2] *= w16^4 = i : folded into the k2 = 2 butterfly below
    v[3 + 4 * 2] = cmul(v[3 + 4 * 2], c_w32[12]);     // w16^6
    v[1 + 4 * 3] = cmul(v[1 + 4 * 3], c_w32[6]);      // w16^3
    v[2 + 4 * 3] = cmul(v[2 + 4 * 3], c_w32[12]);     // w16^6
    v[3 + 4 * 3] = cmul(v[3 + 4 * 3], c_w32[18]);     // w16^9
    ifft4(v[0], v[1], v[2], v[3]);
    ifft4(v[4], v[5], v[6], v[7]);
    ifft4_a2i(v[8], v[9], v[10], v[11]);
    ifft4(v[12], v[13], v[14], v[15]);
    float2 w[16];
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) w[4 * k1 + k2] = v[k1 + 4 * k2];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = w[i];
}

// inverse 32-point DFT of v[0..31] (e^{+2 pi i nk/32}), natural order in and out: one radix-2 stage
// (decimation in frequency: X[2m] = DFT16(a + b)[m], X[2m+1] = DFT16((a - b) w32^n)[m]) and two radix-16
// transforms; 222 packed instructions
__device__ __forceinline__ void ifft32(float2 (&v)[32]) {
    float2 e[16], o[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        e[i] = add2(v[i], v[i + 16]);
        o[i] = sub2(v[i], v[i + 16]);
    }
    // o[n] *= w32^n  (n = 8: i, folded into the first butterfly of ifft16<true>)
#pragma unroll
    for (int n = 1; n < 16; ++n)
        if (n != 8) o[n] = cmul(o[n], c_w32[n]);
    ifft16<false>(e);
    ifft16<true>(o);
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        v[2 * m] = e[m];
        v[2 * m + 1] = o[m];
    }
}

// 16-byte load whose position in the instruction stream is kept (volatile asm)
template <bool SHARED>
__device__ __forceinline__ float4 ld128_pinned(const float4 *p) {
    float4 r;
    if (SHARED) asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(smem_u32(p)));
    else asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// Inverse 1024-point transform by ONE warp.  In: v[i] = x[lane + 32 i].  Out: v[c] = X[lane + 32 c].
//   X[32 k1 + k2] = sum_{n1} w32^{n1 k1} w1024^{n1 k2} sum_{n2} x[n1 + 32 n2] w32^{n2 k2}
// pass A on lane n1 (over n2), twiddle w1024^{lane k2} from the table tw2 (float4 {w^(lane 2j), w^(lane (2j+1))}
// at [j * 32 + lane]; shared or global memory), 32 x 32 transpose through the warp-private tile `ex`
// (EXS float2 per row), pass B on lane k2 (over n1).
// `tile_free()` runs when the warp has read the tile back (it may be overwritten from then on);
// `inputs_read()` runs in front of the first write into the tile, behind a __syncwarp (every lane has
// fetched what it needed from a buffer aliased with the tile).
template <bool TW_SHARED, typename F0, typename F1>
__device__ __forceinline__ void ifft1024(float2 (&v)[32], float2 *ex, const float4 *tw2, int lane, F0 inputs_read, F1 tile_free) {
    {
        // the twiddles are fetched PSH_TWPF ahead of their use: ptxas issues a shared-memory load right in front
        // of its consumer otherwise, and the warp sits out its latency 16 times
        // (measured and rejected: ONE copy of the radix-32 pass in an `unroll 1` loop over the two passes --
        // smaller code, 0.225 against 0.218 ms; 12 warps of 168 registers: 0.230 ms)
        float4 tq[16];
#pragma unroll
        for (int j = 0; j < PSH_TWPF0; ++j) tq[j] = ld128_pinned<TW_SHARED>(tw2 + j * 32 + lane);
        ifft32(v);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (j + PSH_TWPF < 16 && j + PSH_TWPF >= PSH_TWPF0) tq[j + PSH_TWPF] = ld128_pinned<TW_SHARED>(tw2 + (j + PSH_TWPF) * 32 + lane);
            if (j == 0) {
#pragma unroll
                for (int jj = PSH_TWPF0; jj < PSH_TWPF; ++jj) tq[jj] = ld128_pinned<TW_SHARED>(tw2 + jj * 32 + lane);
            }
            const float4 t = tq[j];
            if (j > 0) v[2 * j] = cmul(v[2 * j], make_float2(t.x, t.y));
            v[2 * j + 1] = cmul(v[2 * j + 1], make_float2(t.z, t.w));
        }
        __syncwarp();
        inputs_read();
#pragma unroll
        for (int k2 = 0; k2 < 32; ++k2) ex[k2 * EXS + lane] = v[k2];
        __syncwarp();
#pragma unroll
        for (int n = 0; n < 16; ++n) {
            const float4 t = *reinterpret_cast<const float4 *>(ex + lane * EXS + 2 * n);
            v[2 * n] = make_float2(t.x, t.y);
            v[2 * n + 1] = make_float2(t.z, t.w);
        }
        __syncwarp();
        tile_free();
    }
    ifft32(v);
}

}  // namespace fx3

// atomicAdd by LANE 0 whose result is not needed right away.  ptxas turns an atomic add on a warp-uniform address
// into its warp-aggregated form (vote, one atomic, SHFL of the result) -- also for inline PTX, also under a
// lane predicate -- and the shuffle waits for the round trip (8 % of all warp samples sat there).  An address
// that formally depends on %laneid (0 for the only caller) keeps the plain instruction.
__device__ __forceinline__ unsigned int atom_add_lane0(unsigned int *p, unsigned int v) {
    unsigned int r, l;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
    asm volatile("atom.relaxed.gpu.global.add.u32 %0, [%1], %2;" : "=r"(r) : "l"(p + l), "r"(v));
    return r;
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// debug / test entry: n independent 1024-point transforms, natural order in and out.
// dir +1: inverse (the scan's transform); dir -1: forward through conjugation (the spectra of psh_fft_prepare)
__global__ void __launch_bounds__(128) fft3_debug_kernel(const float2 *__restrict__ in, float2 *out, int n,
                                                         const float4 *__restrict__ tw2, int dir) {
    __shared__ __align__(16) float2 ex[4][32 * fx3::EXS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s = blockIdx.x * 4 + warp;
    if (s >= n) return;
    const float2 *x = in + (size_t)s * fx3::N;
    float2 *y = out + (size_t)s * fx3::N;
    float2 v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        v[i] = x[lane + 32 * i];
        if (dir < 0) v[i].y = -v[i].y;
    }
    fx3::ifft1024<false>(v, ex[warp], tw2, lane, []() {}, []() {});
#pragma unroll
    for (int c = 0; c < 32; ++c) y[lane + 32 * c] = dir < 0 ? make_float2(v[c].x, -v[c].y) : v[c];
}

// spectra of row pairs (1024-sample pieces): one warp per pair; fp32 transform with table twiddles,
// quantised to fp16 pairs with the pair's power-of-two scale; pair norm; MEASURED quantisation error.
__global__ void __launch_bounds__(128) fft3_prep_spectra_kernel(const float *__restrict__ ds, int T, long long row_stride, FftAux a) {
    __shared__ __align__(16) float2 ex[4][32 * fx3::EXS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long pair = (long long)blockIdx.x * 4 + warp;
    if (pair >= a.npairs) return;
    const long long va = 2 * pair, vb = va + 1;
    const bool has_b = vb < a.VR;
    const long long rowa = va / a.nsegv, rowb = (has_b ? vb : va) / a.nsegv;
    const int oa = (int)(va - rowa * a.nsegv) * a.hop, ob = (int)((has_b ? vb : va) - rowb * a.nsegv) * a.hop;
    const float *ya = ds + rowa * row_stride + oa;
    const float *yb = ds + rowb * row_stride + ob;
    float2 v[32];
    double e = 0.0;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const int n = lane + 32 * i;
        const float xa = oa + n < T ? ya[n] : 0.0f;
        const float xb = (ob + n < T && has_b) ? yb[n] : 0.0f;
        v[i] = make_float2(xa, -xb);                       // conj(x): forward = conj(inverse(conj(x)))
        e += (double)xa * (double)xa + (double)xb * (double)xb;
    }
    fx3::ifft1024<false>(v, ex[warp], a.tw2, lane, []() {}, []() {});
    float mx = 0.0f;
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        v[c].y = -v[c].y;
        mx = fmaxf(mx, fmaxf(fabsf(v[c].x), fabsf(v[c].y)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
        e += __shfl_xor_sync(FULL, e, o);
    }
    const float zs = pow2_scale(mx, 14);   // largest component in [2^13, 2^14)
    const double inv = 1.0 / (double)zs;
    uint4 *z4 = reinterpret_cast<uint4 *>(a.Z + (size_t)pair * fx3::N);
    double err = 0.0;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        unsigned int w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = 4 * g + j;
            const float sx = fminf(fmaxf(v[c].x * zs, -65504.0f), 65504.0f);
            const float sy = fminf(fmaxf(v[c].y * zs, -65504.0f), 65504.0f);
            const __half2 h = __floats2half2_rn(sx, sy);
            const float2 f = __half22float2(h);
            const double dx = (double)f.x * inv - (double)v[c].x, dy = (double)f.y * inv - (double)v[c].y;
            err += dx * dx + dy * dy;
            w[j] = *reinterpret_cast<const unsigned int *>(&h);
        }
        z4[g * 32 + lane] = make_uint4(w[0], w[1], w[2], w[3]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) err += __shfl_xor_sync(FULL, err, o);
    if (lane == 0) {
        float4 pi;
        pi.x = __double2float_ru(sqrt(e) * (1.0 + 1e-7));                      // ||y_pair||
        pi.y = __double2float_ru(sqrt(err / (double)fx3::N) * (1.0 + 1e-6));   // ||Z^ - Z||_2 / sqrt(N)
        pi.z = 0.0f;                                                           // es: fft_prep_energy_kernel
        pi.w = zs;                                                             // (replaced by m2 there)
        a.pinfo[pair] = pi;
    }
}

// ------------------------------------------------------------------------------------------
// the scan
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 h2f(unsigned int w) { return __half22float2(*reinterpret_cast<const __half2 *>(&w)); }
// a pair of bf16 (the 1024-point flavour's energies): two integer-pipe instructions, nothing on the fp32 pipe
__device__ __forceinline__ float2 b2f(unsigned int w) { return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }

// The rare part of the epilogue (some window of the warp passed): per-window test, candidate append, upper
// bounds into the threshold histogram.  Register c of a lane holds window lane + 32 c of both rows of the pair.
// Kept SMALL (the warps of an SM run out of phase, so this code is fetched cold): the unrolled part only
// builds the two masks and remembers the last passing window of each row; the appends run in a loop over the
// mask bits.  Only that last window of a lane and row sends its upper bound to the threshold histogram --
// entries may be omitted (k entries at or below a bin still certify its edge), never invented.
__device__ __forceinline__ void fft3_append_candidates(const FftScanParams &p, int b, const float2 (&v)[32], const uint4 *Ys4,
                                                       bool any, float m2, float cu, float rhs, float inv_es, float base0,
                                                       float slack, int pair, int lane, bool count_ub, int ncy) {
    const float INF = __int_as_float(0x7f800000);
    const float2 m22 = make_float2(m2, m2);
    unsigned int ma = 0, mb = 0;
    float ua = INF, ub_ = INF;     // upper bound (scaled units) of the last passing window of row a / row b
    if (any) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            if (g < ncy) {
                const uint4 y4 = Ys4[g * 32 + lane];
                const unsigned int yw[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c = 4 * g + j;
                    const float2 yf = b2f(yw[j]);
                    const float2 val = fx2::fma2(v[c], m22, yf);
                    if (val.x <= rhs && yf.x < INF) { ma |= 1u << c; ua = fmaf(cu, yf.x, val.x); }
                    if (val.y <= rhs && yf.y < INF) { mb |= 1u << c; ub_ = fmaf(cu, yf.y, val.y); }
                }
            }
        }
    }
    const int cnt = __popc(ma) + __popc(mb);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += u;
    }
    const int total = __shfl_sync(FULL, incl, 31);
    if (total == 0) return;
    unsigned int basepos = 0;
    if (lane == 31) basepos = atomicAdd(&p.st[b].ccount, (unsigned int)total);
    basepos = __shfl_sync(FULL, basepos, 31);
    unsigned int pos = basepos + (unsigned int)(incl - cnt);
    unsigned int *dst = p.cand + (size_t)b * p.cap;
    // flat window index of local window 0 of each virtual row: row * T' + piece * hop
    const long long ra = 2 * (long long)pair, rb = ra + 1;
    const unsigned int fa = (unsigned int)((unsigned long long)(ra / p.nsegv) * (unsigned long long)p.Tp
                                           + (unsigned long long)(ra % p.nsegv) * (unsigned long long)p.hop);
    const unsigned int fb = (unsigned int)((unsigned long long)(rb / p.nsegv) * (unsigned long long)p.Tp
                                           + (unsigned long long)(rb % p.nsegv) * (unsigned long long)p.hop);
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
        unsigned int m = h ? mb : ma;
        const unsigned int f0 = (h ? fb : fa) + (unsigned int)lane;
        while (m != 0u) {
            const int c = __ffs(m) - 1;
            m &= m - 1u;
            if (pos < p.cap) dst[pos] = f0 + 32u * (unsigned int)c;
            ++pos;
        }
        const float u0 = h ? ub_ : ua;
        if (count_ub && u0 < INF) {   // upper bound of the window's exact squared distance
            const float ub = fmaxf((((u0 + 5.9604644775390625e-8f) * inv_es + base0) + 2.0f * slack) * 1.000001f, 0.0f);
            const int bin = (int)(__float_as_uint(ub) >> 13) - hist_base(p.st[b].q2);
            if (ub < INF && bin < HB) hist_add_ub(p.hist + (size_t)b * HSTRIDE, bin < 0 ? 0 : bin, 1u);
        }
    }
}

// End of a WARP's seeding pass: arrive (every histogram increment of this warp is ordered before the arrival:
// __syncwarp + fence); the warp whose arrival is the `seed_need`-th derives the thresholds of all queries and
// publishes them, every other warp waits for that (bounded: arrivals never wait for anybody) and picks them up.
// Only ONE warp per CTA polls the global flag (the first of the CTA to get here) and relays it through shared
// memory: 2368 warps polling one L2 line every 200 ns kept the line so busy that the arrivals and the
// publication themselves queued behind the polls (measured: 12 us from the median arrival to the release).
// s_seed[0]: ticket of the CTA's poller, s_seed[1]: the launch's thresholds are published.
// VIF (one query): the flag word CARRIES the threshold (its float bits, never 0), so the publication needs no
// fence between threshold and flag and the waiting warps no load of the published threshold: two L2 round
// trips less between the last arrival and the release.
template <bool VIF>
__device__ __forceinline__ void fft3_seed_rendezvous(const FftScanParams &p, const float *s_q2, float *s_thr, int lane,
                                                     volatile unsigned int *s_seed) {
    volatile unsigned int *done = p.hist + H_DONE;
    unsigned int *relay = const_cast<unsigned int *>(&s_seed[1]);
    __syncwarp();
    int last = 0;
    unsigned int seen = 0u;
    if (lane == 0) {
        __threadfence();
        const unsigned int ticket = atomicAdd(p.hist + H_ARR, 1u);
        last = (ticket + 1u == p.seed_need) ? 1 : 0;
        if (!last) {
            const bool poller = atomicAdd(const_cast<unsigned int *>(&s_seed[0]), 1u) == 0u;
            const unsigned long long t0 = globaltimer_ns();
            // (the relay flag is read and written with shared-memory atomics: a deliberate flag hand-off, and
            // compute-sanitizer's racecheck stays clean)
            while ((seen = atomicOr(relay, 0u)) == 0u) {
                if (poller) {
                    seen = *done;
                    if (seen != 0u) { atomicExch(relay, seen); break; }
                }
                if (globaltimer_ns() - t0 > 2000000ull) break;   // 2 ms: thresholds stay loose, the call re-runs safely
                __nanosleep(poller ? 200 : 100);
            }
            if (VIF && seen != 0u) atomicMin(reinterpret_cast<unsigned int *>(&s_thr[0]), seen);
        }
    }
    last = __shfl_sync(FULL, last, 0);
    if (last) {
        for (int b = 0; b < p.nq; ++b) fft_refresh_threshold(p, b, s_q2[b], s_thr);
        __syncwarp();
        if (lane == 0) {
            unsigned int word = 1u;
            if (VIF) word = __float_as_uint(s_thr[0]);   // (positive float or +inf: never 0; lane 0 merged it itself)
            else __threadfence();
            *done = word;
            atomicExch(relay, word);
        }
    } else if (!VIF || __shfl_sync(FULL, seen, 0) == 0u) {   // (a group of queries, or the wait timed out)
        if (lane < p.nq) {
            const unsigned int tb = *reinterpret_cast<volatile unsigned int *>(p.hist + (size_t)lane * HSTRIDE + H_THR);
            atomicMin(reinterpret_cast<unsigned int *>(&s_thr[lane]), tb);
        }
    }
    __syncwarp();
}

// One WARP per row pair and iteration: Z^ * conj(Q)/N -> inverse 1024-point FFT -> (D_a[t], D_b[t]); the
// lower bound, thresholds, seeding and candidate lists of fft_scan_kernel (see there), at warp granularity:
// every warp draws its own pairs (atomic slot counter, permuted order), owns two mbarriers and its staging
// buffers (TMA bulk copies issued one pair ahead by lane 0), arrives at the seeding rendezvous on its own
// and takes its turn re-deriving the thresholds.  The CTA (16 warps, one per SM; 13 for a group of queries)
// only shares the twiddle table, the thresholds and -- one query -- the query's spectrum.
// One query: the pair's spectrum is staged INSIDE the transpose tile (it is in registers before the tile
// is written) and the next one is fetched when the tile has been read back.  A group of queries keeps the
// spectrum in a buffer of its own for all its transforms.
// NCY > 0: the number of 128-window groups of a piece's energy row is a compile-time constant (7 for
// 129 <= W <= 253 on trajectories longer than one piece: the epilogue's eight guarded groups cost ~40
// instructions of control per transform otherwise); NCY = 0: taken from p.ncy.
template <bool EMB, bool SINGLE, int NCY>
__global__ void __launch_bounds__(fx3::WARPS_SINGLE * 32, 1) fft_scan_warp_kernel(const FftScanParams p) {
    extern __shared__ __align__(128) unsigned char fsm[];
    __shared__ float s_thr[QG_MAX], s_q2[QG_MAX], s_qmax[QG_MAX], s_gn[QG_MAX];
    __shared__ unsigned int s_seed[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    constexpr bool single = SINGLE;   // one query: its spectrum in shared memory, the pair's spectrum staged inside the tile
    const int ncy = NCY > 0 ? NCY : p.ncy;
    const int nq = SINGLE ? 1 : p.nq;
    const float INF = __int_as_float(0x7f800000);
    float4 *tw2s = reinterpret_cast<float4 *>(fsm);
    const float4 *Qs4 = reinterpret_cast<const float4 *>(fsm + fx3::TW_BYTES);
    unsigned char *wbase = fsm + fx3::TW_BYTES + (single ? fx3::Q_BYTES + warp * fx3::WARP_BYTES_ALIAS : warp * fx3::WARP_BYTES_SEP);
    float2 *ex = reinterpret_cast<float2 *>(wbase);
    unsigned char *zbuf = single ? wbase : wbase + fx3::EX_BYTES;
    unsigned char *ybuf = zbuf + (single ? fx3::EX_BYTES : fx3::Z_BYTES);
    const uint4 *Zs4 = reinterpret_cast<const uint4 *>(zbuf);
    const uint4 *Ys4 = reinterpret_cast<const uint4 *>(ybuf);
    const float4 *pis = reinterpret_cast<const float4 *>(ybuf + fx3::Y_BYTES);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(ybuf + fx3::Y_BYTES + 16);
    const uint32_t barZ = smem_u32(&bars[0]), barY = smem_u32(&bars[1]);

    for (int i = tid; i < fx3::TW_BYTES / 16; i += blockDim.x) tw2s[i] = __ldg(p.tw2 + i);
    if (single)
        for (int i = tid; i < fx3::Q_BYTES / 16; i += blockDim.x)
            reinterpret_cast<float4 *>(fsm + fx3::TW_BYTES)[i] = __ldg(reinterpret_cast<const float4 *>(p.Qc) + i);
    if (tid < nq) {
        const float t0 = ld_volatile_f32(&p.st[tid].thr_fast);
        s_thr[tid] = EMB ? t0 * p.thr_widen : t0;
        s_q2[tid] = p.st[tid].q2;
        s_gn[tid] = p.st[tid].gnorm;
        s_qmax[tid] = 0.0f;
    }
    if (lane == 0) {
        mbar_init(barZ, 1);
        mbar_init(barY, 1);
        mbar_fence_init();
    }
    if (tid < 2) s_seed[tid] = 0u;
    __syncthreads();
    for (int b = 0; b < nq; ++b) {   // max_k |FFT(q)_k| from the partial maxima (positive floats order as uints)
        float m = 0.0f;
        for (int i = tid; i < p.nqmax; i += blockDim.x) m = fmaxf(m, __ldg(p.qmaxp + (size_t)b * QMAXP + i));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
        if (lane == 0) atomicMax(reinterpret_cast<unsigned int *>(&s_qmax[b]), __float_as_uint(m));
    }
    __syncthreads();   // the last CTA barrier: from here on every warp is on its own

    const int gw = (int)blockIdx.x * nw + warp, tw = (int)gridDim.x * nw;
    const int slot = p.i0 + gw;
    if (slot >= p.i1) return;   // (a seeding launch is sized so that every warp owns a pair)
    // optional per-warp timeline (8 globaltimer stamps per warp; tests/timeline.py), NULL in production
#define PSH_STAMP3(i) do { if (p.dbg != nullptr && lane == 0) p.dbg[(size_t)gw * 8 + (i)] = globaltimer_ns(); } while (0)
    PSH_STAMP3(0);

    const uint32_t ybytes = (uint32_t)ncy * 512u;
    auto issue_z = [&](int pr) {  // lane 0
        mbar_expect_tx(barZ, (uint32_t)fx3::Z_BYTES);
        bulk_g2s(smem_u32(zbuf), p.Z + (size_t)pr * fx3::N, (uint32_t)fx3::Z_BYTES, barZ);
    };
    auto issue_y = [&](int pr) {  // lane 0: the energy rows and the pair's statistics
        mbar_expect_tx(barY, ybytes + (uint32_t)sizeof(float4));
        bulk_g2s(smem_u32(ybuf), p.Y2 + (size_t)pr * fx3::N, ybytes, barY);
        bulk_g2s(smem_u32(pis), p.pinfo + pr, (uint32_t)sizeof(float4), barY);
    };

    int pair = fft_pair_of_slot(p, slot);
    if (lane == 0) { issue_z(pair); issue_y(pair); }

    uint32_t phZ = 0, phY = 0;
    bool seeding = p.seed != 0;
    bool staged = false;          // this pair's spectrum and energies are already in shared memory (a group's pass after seeding)
    const bool rerun = nq > 1;  // seeding a group of queries: the pair is transformed twice
    const float cu = p.ub_y_coef;
    for (int iter = 0;;) {
        // lane 0 draws the slot behind this pair now; the atomic's round trip hides behind the first pass
        const bool draw = !(seeding && rerun);
        unsigned int drawn = 0;
        if (draw && lane == 0) drawn = atom_add_lane0(p.hist + H_SLOT, 1u);
        // thresholds other warps have published: fetched now, merged at the end of the iteration
        // (measured and rejected: picking the thresholds up every iteration / re-deriving them more often during the
        // first iterations after seeding, +0.4 % and +1.2 %)
        const bool pick = ((iter + gw) & 3) == 0;
        unsigned int pub = 0x7f800000u;
        if (pick && lane < nq) pub = __ldcg(p.hist + (size_t)lane * HSTRIDE + H_THR);
        if (!staged) { mbar_wait(barZ, phZ); phZ ^= 1; }
        int npair = -1;
        for (int b = 0; b < nq; ++b) {
            float2 v[32];
            {
                const float4 *Q4 = single ? Qs4 : reinterpret_cast<const float4 *>(p.Qc + (size_t)b * fftx::N);
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const uint4 z4 = Zs4[g * 32 + lane];
                    float4 qa, qb;
                    if (single) { qa = Q4[(2 * g) * 32 + lane]; qb = Q4[(2 * g + 1) * 32 + lane]; }
                    else { qa = __ldg(Q4 + (2 * g) * 32 + lane); qb = __ldg(Q4 + (2 * g + 1) * 32 + lane); }
                    v[4 * g + 0] = fx2::cmul(h2f(z4.x), make_float2(qa.x, qa.y));
                    v[4 * g + 1] = fx2::cmul(h2f(z4.y), make_float2(qa.z, qa.w));
                    v[4 * g + 2] = fx2::cmul(h2f(z4.z), make_float2(qb.x, qb.y));
                    v[4 * g + 3] = fx2::cmul(h2f(z4.w), make_float2(qb.z, qb.w));
                }
            }
            const bool last_q = b == nq - 1;
            // the pair behind this one: lane 0 turns the slot it drew at the top into a pair when the first pass
            // has hidden the atomic's round trip, and issues the copy of its spectrum
            auto next_pair = [&]() {
                int np = -1;
                if (lane == 0) {
                    const long long ns = (long long)p.i0 + (long long)tw + (long long)drawn;
                    np = (draw && ns < (long long)p.i1) ? fft_pair_of_slot(p, (int)ns) : -1;
                }
                return np;
            };
            int np0 = -1;
            fx3::ifft1024<true>(v, ex, tw2s, lane,
                [&]() {   // every lane holds its part of the staged spectrum
                    if (!single && last_q) {
                        np0 = next_pair();
                        if (lane == 0 && np0 >= 0) issue_z(np0);
                    }
                },
                [&]() {   // the tile (= the spectrum's staging buffer) has been read back
                    if (single) {
                        np0 = next_pair();
                        if (lane == 0 && np0 >= 0) { fence_proxy_async_smem(); issue_z(np0); }
                    }
                });
            if (last_q) npair = __shfl_sync(FULL, np0, 0);
            if (b == 0 && !staged) { mbar_wait(barY, phY); phY ^= 1; }  // the pair's window energies have landed
            const float4 pi = *pis;
            const float yn = pi.x, zq = pi.y, es = pi.z, m2 = pi.w;
            const float inv_es = 1.0f / es;               // es is a power of two
            const float q2 = s_q2[b], qmax = s_qmax[b], gn = s_gn[b];
            float slack;
            if (EMB) slack = (2.0f * p.cf_u * qmax * yn + 2.0f * zq * gn + p.slack_coef * q2 + p.g_coef * gn * yn) * 1.0001f;
            else slack = (2.0f * p.cf_u * qmax * yn + 2.0f * zq * gn + 7.152557373046875e-7f * (q2 + yn * yn)) * 1.0001f;
            const float base0 = q2 - slack;   // LB = (Y2^ - 2 D^) + base0, kept iff LB <= thr
            const float2 m22 = make_float2(m2, m2);
            if (seeding) {
                // minimum over the lane's windows of the UPPER bound (scaled units); windows beyond T' are +inf
                float mn = INF;
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    if (g < ncy) {
                        const uint4 y4 = Ys4[g * 32 + lane];
                        const unsigned int yw[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 yf = b2f(yw[j]);
                            const float2 ub2 = fx2::fma2(yf, make_float2(cu, cu), fx2::fma2(v[4 * g + j], m22, yf));
                            mn = fminf(mn, fminf(ub2.x, ub2.y));
                        }
                    }
                }
                // one entry per group of `seed_group` lanes (a power of two): the atomics of all warps land in a few
                // dozen bins and serialise in L2 -- 32 entries per warp kept every warp ~30 us at the rendezvous
                for (int o = 1; o < p.seed_group; o <<= 1) mn = fminf(mn, __shfl_xor_sync(FULL, mn, o));
                // true units; 2^-24: an energy in fp16's subnormal range was floored by at most that much
                const float ub = fmaxf((((mn + 5.9604644775390625e-8f) * inv_es + base0) + 2.0f * slack) * 1.000001f, 0.0f);
                int bin = (int)(__float_as_uint(ub) >> 13) - hist_base(q2);
                bin = bin < 0 ? 0 : bin;
                const bool cnt = ub < INF && bin < HB && (lane & (p.seed_group - 1)) == 0;   // false for +inf (no valid window) and NaN
                // lanes whose minima share a bin send ONE increment
                const unsigned int same = __match_any_sync(FULL, cnt ? bin : HB + lane);
                if (cnt && lane == __ffs(same) - 1) hist_add_ub(p.hist + (size_t)b * HSTRIDE, bin, (unsigned int)__popc(same));
                if (rerun) continue;
                PSH_STAMP3(1);
                fft3_seed_rendezvous<SINGLE>(p, s_q2, s_thr, lane, s_seed);
                PSH_STAMP3(2);
            }
            const float thr = s_thr[b];
            const float tdiff = thr - base0;
            const float rhs = fmaf(fabsf(tdiff), 9.5367431640625e-7f, tdiff) * es;   // (+2^-20: roundings of rhs and of the fma below)
            float mn = INF;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                if (g < ncy) {
                    const uint4 y4 = Ys4[g * 32 + lane];
                    const unsigned int yw[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 val = fx2::fma2(v[4 * g + j], m22, b2f(yw[j]));
                        mn = fminf(mn, fminf(val.x, val.y));
                    }
                }
            }
            const bool any = mn <= rhs;   // (+inf <= +inf while thr = +inf: sorted out per window in the append)
            if (__any_sync(FULL, any))    // rare
                fft3_append_candidates(p, b, v, Ys4, any, m2, cu, rhs, inv_es, base0, slack, pair, lane, !(p.seed != 0 && iter == 0), ncy);
        }
        if (seeding && rerun) {
            // a group of queries: arrive, wait for the thresholds, then the same pair again
            fft3_seed_rendezvous<false>(p, s_q2, s_thr, lane, s_seed);
            seeding = false;
            staged = true;
            continue;
        }
        seeding = false;
        staged = false;
        __syncwarp();   // every lane is done with the staged energies
        if (lane == 0 && npair >= 0) issue_y(npair);
        if (pick && lane < nq && pub < __float_as_uint(s_thr[lane])) atomicMin(reinterpret_cast<unsigned int *>(&s_thr[lane]), pub);
        if (iter == 0) PSH_STAMP3(3);
        if (iter == 1) PSH_STAMP3(4);
        if (iter == 8) PSH_STAMP3(5);
        if (iter == 24) PSH_STAMP3(6);
        if (npair < 0) { PSH_STAMP3(7); break; }
        if (((iter + gw) & p.refresh_mask) == 0)
            for (int b = 0; b < nq; ++b) fft_refresh_threshold(p, b, s_q2[b], s_thr);
        pair = npair;
        ++iter;
    }
}
