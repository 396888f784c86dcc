// pshadow.cu -- B200 (sm_100a) path-shadowing scan: kernels + the C ABI of include/pshadow.h.
//
// Replaces, for Identity + RelativeMSE + PredictionContext, the reference hot path
//   PathShadowing.batched_distance   path_shadowing.py:97-179
//   PathShadowing.shadow (gather)    path_shadowing.py:210-216
//   predict_from_paths               path_shadowing.py:234-254 (+ statistics.py:5-16)
// The reference materialises all R*T' windows (conv1d with eye(W)), subtracts, norms, divides,
// top-k's per split and merges.  Here the ensemble rows are streamed once per query group:
// each warp stages a row segment in shared memory with a TMA bulk copy (cp.async.bulk +
// mbarrier, double buffered), every lane slides WPT consecutive windows through registers, and
// windows that beat the running threshold are appended to a per-query candidate list from
// which a radix select keeps the k best.  The distance arithmetic reproduces the reference's
// CPU bit pattern (sequential, non-fused fp32), so indices and distances are bit-identical.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <math.h>
#include <string.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "pshadow.h"

namespace {

// ------------------------------------------------------------------------------------------
// constants
// ------------------------------------------------------------------------------------------
constexpr int WPT = 12;              // consecutive windows per lane (48-byte lane stride: LDS.128 conflict-free)
constexpr int SEG = 32 * WPT;        // windows per warp task
constexpr int RING = 16;             // register ring of staged samples (WPT + one float4 in flight)
constexpr int SCAN_WARPS = 8;        // warps per scan CTA
constexpr int SCAN_THREADS = SCAN_WARPS * 32;
constexpr int SEL_THREADS = 1024;
constexpr int SORT_SMEM_MAX = 16384; // keys sorted in shared memory (128 KiB); larger k sorts in global
constexpr int QG_MAX = 32;           // queries staged per scan launch
constexpr unsigned FULL = 0xffffffffu;

std::atomic<uint64_t> g_launches{0};

struct QState {
    unsigned long long tau_key;  // inclusive threshold on (distance bits << 32 | flat window index)
    float s_thr;                 // largest squared numerator whose distance is <= tau's distance
    float qnorm;                 // ||q|| in the reference's 8-lane order
    unsigned int count;          // keys appended to the current buffer (may exceed cap: overflow)
    unsigned int overflow;
    unsigned int cur;            // current ping-pong buffer
    unsigned int ccount;         // filter candidates awaiting the exact re-rank
    float thr_fast;              // s_thr widened by the exact sequence's own rounding (filter compare)
    float q2;                    // sum q_j^2 (fp64 accumulated, rounded once)
    float qmax;                  // max_k |FFT(q)_k| (fft flavour), rounded up
    unsigned int sticky;         // overflow of ANY scan since the last psh_scan_overflowed(): survives qprep,
    unsigned int magic;          // valid only while magic == QSTATE_MAGIC (the workspace starts as garbage)
    float gnorm;                 // ||g||_2 of the correlated vector (fft flavour of the embedded scan), rounded up
    unsigned int pad[2];
};
constexpr unsigned int QSTATE_MAGIC = 0x50534831u;
static_assert(sizeof(QState) == 64, "QState layout");

struct ScanParams {
    const float *ds;
    long long row_stride;
    int T, Tp, W, nseg;
    unsigned int nseg_m, nseg_s;  // multiply-high division by nseg
    unsigned int ntasks;   // (i1 - i0) * nseg, < 2^32
    long long R;
    double inv_R;
    long long i0, i1;      // range of permuted row slots scanned by this launch
    long long perm;        // row = (slot * perm) % R, gcd(perm, R) = 1
    const float *queries;  // (nq, W)
    int nq;
    QState *st;            // (nq)
    unsigned long long *keys;  // (nq, 2, cap)
    unsigned int *cand;        // (nq, cap) flat window indices that passed the filter
    unsigned int cap;
    int bulk_ok;           // rows are 16-byte aligned: TMA bulk staging
    int buf_floats;        // floats per staging buffer (multiple of 4)
    int wpad;              // padded query stride in shared memory
    int epl;               // filter: samples per lane of the squared-prefix pass (odd multiple of 4)
    int pfx_floats;        // filter: floats of the per-warp prefix buffer
    float cw;              // filter: slack coefficient (W + 256) * 2^-24
    int pair_mode;         // fft flavour: slots enumerate the two members of permuted PAIRS of virtual rows
    long long npairs;
    double inv_np;
    // virtual rows (fft flavour): a trajectory longer than one 4096-point transform is cut into nsegv
    // overlapping pieces; virtual row v = (row v / nsegv, piece v % nsegv) owns the windows
    // [piece*hop, piece*hop + span) of its row.  nsegv = 1, span = T' when T <= 4096.
    int nsegv, hop, span;
    long long VR;
};

// ------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA bulk copy (SASS: SYNCS / UBLKCP)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "PSH_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra PSH_DONE;\n"
        "bra PSH_WAIT;\n"
        "PSH_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}

__device__ __forceinline__ void cp_async_4(uint32_t dst, const float *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.wait_all;" ::: "memory");
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void *src) {   // L2 only (.cg): never a stale L1 line
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ float ld_volatile_f32(const float *p) {
    return *reinterpret_cast<const volatile float *>(p);
}
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) {
    return *reinterpret_cast<const volatile unsigned long long *>(p);
}

// the reference's distance from its squared numerator: sqrt then IEEE divide (path_distance.py:65)
__device__ __forceinline__ float dist_from_s(float s, float qn) { return __fdiv_rn(__fsqrt_rn(s), qn); }

#include "pshadow_fft.cuh"
#include "pshadow_fft2.cuh"

// ------------------------------------------------------------------------------------------
// query preparation: ||q|| in torch's contiguous-reduction order (8 interleaved partial sums,
// lanes added 0..7, scalar tail), state reset.  path_distance.py:65 `x.norm(dim=-1)`.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) qprep_kernel(const float *__restrict__ q, int W, int nq, QState *st) {
    // one warp per query; lanes 0..7 own torch's 8 interleaved partial sums
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= nq) return;
    const float *x = q + (size_t)b * W;
    const int n8 = (W / 8) * 8;
    float acc = 0.0f;
    if (lane < 8)
        for (int j = lane; j < n8; j += 8) acc = __fadd_rn(acc, __fmul_rn(x[j], x[j]));
    float s = 0.0f;
#pragma unroll
    for (int l = 0; l < 8; ++l) s = __fadd_rn(s, __shfl_sync(FULL, acc, l));
    for (int j = n8; j < W; ++j) s = __fadd_rn(s, __fmul_rn(x[j], x[j]));  // scalar tail (all lanes alike)
    double q2 = 0.0;
    for (int j = lane; j < W; j += 32) q2 += (double)x[j] * (double)x[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q2 += __shfl_xor_sync(FULL, q2, o);
    if (lane == 0) {
        QState z;
        z.sticky = (st[b].magic == QSTATE_MAGIC) ? st[b].sticky : 0u;
        z.magic = QSTATE_MAGIC;
        z.tau_key = ~0ull;
        z.s_thr = __int_as_float(0x7f800000);
        z.qnorm = __fsqrt_rn(s);
        z.count = 0;
        z.overflow = 0;
        z.cur = 0;
        z.ccount = 0;
        z.thr_fast = __int_as_float(0x7f800000);
        z.q2 = (float)q2;
        z.qmax = 0.0f;
        z.gnorm = 0.0f;
        for (int i = 0; i < 2; ++i) z.pad[i] = 0;
        st[b] = z;
    }
}

// ------------------------------------------------------------------------------------------
// the scan kernel (exact and filter flavours share staging, task decode and register tiling)
// ------------------------------------------------------------------------------------------
// Work unit ("task") = SEG consecutive windows of one trajectory, owned by one warp:
//   * lane 0 stages the SEG+W-1 samples the task touches with ONE TMA bulk copy
//     (cp.async.bulk -> mbarrier complete_tx), double buffered per warp: the copy of the warp's
//     next task is in flight while the current one is evaluated; no CTA-wide barrier exists in
//     the steady state;
//   * each lane owns WPT consecutive windows and slides them through a 16-register ring: one
//     LDS.128 of samples and one LDS.128 (broadcast) of query values feed 4 reduction steps x
//     WPT windows of arithmetic;
//   * exact flavour: acc = fl(acc + fl(fl(q_j - y)^2)), j ascending, never FMA -- the
//     reference's reduction order (path_distance.py:65 on the conv1d windows), hence its bits;
//   * filter flavour: one FMA per element, see below.
template <bool EXACT, bool GUARD>
__device__ __forceinline__ void step_block(float (&acc)[WPT], float (&ring)[RING], const float *__restrict__ yb,
                                           const float *__restrict__ qs, int rem) {
    float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int jj = 0; jj < RING; ++jj) {
        if (GUARD && jj >= rem) break;
        if ((jj & 3) == 0) {
            const float4 v = *reinterpret_cast<const float4 *>(yb + jj + WPT);
            ring[(jj + WPT + 0) & (RING - 1)] = v.x;
            ring[(jj + WPT + 1) & (RING - 1)] = v.y;
            ring[(jj + WPT + 2) & (RING - 1)] = v.z;
            ring[(jj + WPT + 3) & (RING - 1)] = v.w;
            q4 = *reinterpret_cast<const float4 *>(qs + jj);
        }
        const float qj = (jj & 3) == 0 ? q4.x : (jj & 3) == 1 ? q4.y : (jj & 3) == 2 ? q4.z : q4.w;
#pragma unroll
        for (int w = 0; w < WPT; ++w) {
            if (EXACT) {
                const float df = __fsub_rn(qj, ring[(jj + w) & (RING - 1)]);
                acc[w] = __fadd_rn(acc[w], __fmul_rn(df, df));
            } else {
                acc[w] = fmaf(qj, ring[(jj + w) & (RING - 1)], acc[w]);
            }
        }
    }
}

// Note on the filter flavour's inner loop (measured, profiles/r01_*): a scalar FFMA reads three
// registers while the register file serves one even and one odd register per cycle; the query
// value sits in the operand-reuse cache, but with a sliding ring every accumulator meets every
// sample slot, so ptxas cannot keep (sample, accumulator) in opposite banks everywhere: ~60 % of
// the FFMAs take two cycles (66 % FP32-pipe utilisation).  Two alternatives were built and
// measured slower on B200: a dual-ring layout (samples duplicated with flipped parity; ptxas
// re-schedules across steps and loses the property) and packed FFMA2 (fma.rn.f32x2; 1.67 ms vs
// 1.44 ms per query -- three 64-bit operand reads per instruction without reuse).
struct Task { long long row; int t0; int nvalid; int tp_eff; };

__device__ __forceinline__ Task decode_task(const ScanParams &p, unsigned int task) {
    // task -> (row slot, segment): division by the launch-invariant nseg via multiply-high
    const unsigned int sr = (unsigned int)(((unsigned long long)__umulhi(task, p.nseg_m) + task) >> p.nseg_s);
    const unsigned int seg = task - sr * (unsigned int)p.nseg;
    // row = (slot * perm) mod R: quotient estimated in fp64 (exact to +-1), remainder fixed up
    const unsigned long long slot = (unsigned long long)(p.i0 + sr);
    const unsigned long long modn = p.pair_mode ? (unsigned long long)p.npairs : (unsigned long long)p.R;
    const unsigned long long prod = (p.pair_mode ? (slot >> 1) : slot) * (unsigned long long)p.perm;
    const unsigned long long q = __double2ull_rz(__ull2double_rz(prod) * (p.pair_mode ? p.inv_np : p.inv_R));
    long long r = (long long)(prod - q * modn);
    if (r < 0) r += (long long)modn;
    else if (r >= (long long)modn) r -= (long long)modn;
    Task t;
    if (p.pair_mode) {
        const long long v = 2 * r + (long long)(slot & 1ull);       // virtual row
        const long long row = v / p.nsegv;
        const int piece = (int)(v - row * p.nsegv);
        t.row = v < p.VR ? row : p.R - 1;
        t.t0 = piece * p.hop + (int)seg * SEG;
        t.tp_eff = v < p.VR ? min(p.Tp, piece * p.hop + p.span) : 0;  // phantom partner of an odd count
        if (t.t0 >= t.tp_eff) { t.t0 = 0; t.tp_eff = 0; }           // nothing of this piece left to scan
    } else {
        t.row = r;
        t.t0 = (int)seg * SEG;
        t.tp_eff = p.Tp;
    }
    t.nvalid = min(SEG + p.W - 1, p.T - t.t0);
    return t;
}

// Filter flavour (PSH_MODE_FILTER): ||q - y_t||^2 = Q2 + Y2_t - 2 D_t with D_t = sum_j q_j y_{t+j}
// (one FFMA chain per window) and Y2_t = sum_j y_{t+j}^2 taken from a per-task prefix sum of
// squares (one warp scan).  The kernel evaluates a rigorous LOWER BOUND of the true squared
// distance,
//     LB = Q2 + Y2^ - 2 D^ - cw (Q2 + Ptot),      cw = (W + 256) 2^-24,
// (D^: |D^-D| <= gamma_W sum|q_j y_j| <= gamma_W (Q2+Y2)/2;  Y2^: difference of two prefix
// values of depth <= 46 roundings, |Y2^-Y2| <= 96u Ptot;  Q2 rounded once;  <= 16u (Q2+Ptot) for
// the combination itself;  Ptot = sum of squares of the whole staged segment >= every Y2_t)
// and appends the window to the query's candidate list iff LB <= thr_fast, where thr_fast is
// s_thr widened by the exact sequence's own worst-case rounding (select_kernel).  Every window
// whose EXACT distance beats the threshold therefore passes; the survivors (a few 1e-4 of all
// windows) are re-evaluated with the reference's exact sequence by rerank_kernel, so the final
// top-k is bit-identical to PSH_MODE_EXACT.  NaN/Inf bounds pass (decided exactly later).
template <bool EXACT>
__global__ void __launch_bounds__(SCAN_THREADS, EXACT ? 2 : 3) scan_kernel(const ScanParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *qs_all = reinterpret_cast<float *>(smem_raw);
    float *bufs = qs_all + (size_t)p.nq * p.wpad;
    float *pfx_all = bufs + (size_t)SCAN_WARPS * 2 * p.buf_floats;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(
        pfx_all + (EXACT ? (size_t)0 : (size_t)SCAN_WARPS * p.pfx_floats));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *mybuf = bufs + (size_t)warp * 2 * p.buf_floats;
    float *pfx = pfx_all + (size_t)warp * p.pfx_floats;  // pfx[4 + i] = sum_{e<=i} y_e^2, pfx[0..3] = 0
    const uint32_t bar0 = smem_u32(&bars[warp * 2]);

    // stage the queries (zero padded to wpad so the float4 fetch of the tail stays in bounds)
    for (int i = threadIdx.x; i < p.nq * p.wpad; i += SCAN_THREADS) {
        const int b = i / p.wpad, j = i - b * p.wpad;
        qs_all[i] = j < p.W ? p.queries[(size_t)b * p.W + j] : 0.0f;
    }
    if (!EXACT && lane < 4) pfx[lane] = 0.0f;
    if (p.bulk_ok && lane == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        mbar_fence_init();
    }
    __syncthreads();

    const unsigned int ntasks = p.ntasks;
    const unsigned int gw = blockIdx.x * SCAN_WARPS + warp;
    const unsigned int nw = gridDim.x * SCAN_WARPS;
    if (gw >= ntasks) return;

    auto issue = [&](const Task &t, int which) {  // lane 0 only
        const uint32_t bytes = (uint32_t)((t.nvalid + 3) & ~3) * 4u;
        const uint32_t bar = bar0 + 8u * which;
        mbar_expect_tx(bar, bytes);
        bulk_g2s(smem_u32(mybuf + (size_t)which * p.buf_floats), p.ds + t.row * p.row_stride + t.t0, bytes, bar);
    };

    Task tk = decode_task(p, gw);
    uint32_t phase0 = 0, phase1 = 0;
    if (p.bulk_ok && lane == 0) issue(tk, 0);

    int n = 0;
    for (unsigned int task = gw; task < ntasks; ++n) {
        const int cur = n & 1;
        const unsigned int next = task + nw;
        const bool have_next = next < ntasks && next > task;
        Task tn = tk;
        if (have_next) tn = decode_task(p, next);
        float *buf = mybuf + (size_t)cur * p.buf_floats;
        if (p.bulk_ok) {
            if (have_next && lane == 0) issue(tn, cur ^ 1);
            if (cur == 0) { mbar_wait(bar0, phase0); phase0 ^= 1; }
            else { mbar_wait(bar0 + 8, phase1); phase1 ^= 1; }
        } else {
            const float *src = p.ds + tk.row * p.row_stride + tk.t0;
            for (int i = lane; i < tk.nvalid; i += 32) buf[i] = __ldg(src + i);
        }
        const int t0 = tk.t0;
        const float *yb = buf + lane * WPT;
        const int tl = t0 + lane * WPT;  // first window of this lane
        const unsigned int flat0 =
            (unsigned int)((unsigned long long)tk.row * (unsigned long long)p.Tp + (unsigned long long)tl);
        const bool all_valid = t0 + SEG <= tk.tp_eff;

        float y2[WPT];
        float ptot = 0.0f;
        if (!EXACT) {
            // samples beyond the valid part of the row count as zeros in the prefix sums
            for (int i = tk.nvalid + lane; i < 32 * p.epl; i += 32) buf[i] = 0.0f;
            __syncwarp();
            // ---- inclusive prefix sums of squares over the staged segment ----
            const int e0 = lane * p.epl;
            const float *yl = buf + e0;
            float tot = 0.0f;
            for (int i = 0; i < p.epl; i += 4) {
                const float4 v = *reinterpret_cast<const float4 *>(yl + i);
                tot = fmaf(v.x, v.x, tot); tot = fmaf(v.y, v.y, tot);
                tot = fmaf(v.z, v.z, tot); tot = fmaf(v.w, v.w, tot);
            }
            float incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float v = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += v;
            }
            ptot = __shfl_sync(FULL, incl, 31);
            float run = incl - tot;  // exclusive offset of this lane
            for (int i = 0; i < p.epl; i += 4) {
                const float4 v = *reinterpret_cast<const float4 *>(yl + i);
                float4 o4;
                run = fmaf(v.x, v.x, run); o4.x = run;
                run = fmaf(v.y, v.y, run); o4.y = run;
                run = fmaf(v.z, v.z, run); o4.z = run;
                run = fmaf(v.w, v.w, run); o4.w = run;
                *reinterpret_cast<float4 *>(pfx + 4 + e0 + i) = o4;
            }
            __syncwarp();
            // Y2 of this lane's windows: pfx[4 + t+W-1] - pfx[4 + t-1]
            const int tloc = lane * WPT;
            float lo[16];
            const float4 a = *reinterpret_cast<const float4 *>(pfx + tloc + 0);
            const float4 c = *reinterpret_cast<const float4 *>(pfx + tloc + 4);
            const float4 e = *reinterpret_cast<const float4 *>(pfx + tloc + 8);
            const float4 g = *reinterpret_cast<const float4 *>(pfx + tloc + 12);
            lo[0] = a.x; lo[1] = a.y; lo[2] = a.z; lo[3] = a.w; lo[4] = c.x; lo[5] = c.y; lo[6] = c.z; lo[7] = c.w;
            lo[8] = e.x; lo[9] = e.y; lo[10] = e.z; lo[11] = e.w; lo[12] = g.x; lo[13] = g.y; lo[14] = g.z; lo[15] = g.w;
            const float *hi = pfx + tloc + p.W + 3;
#pragma unroll
            for (int w = 0; w < WPT; ++w) y2[w] = hi[w] - lo[w + 3];
        } else {
            if (!p.bulk_ok) __syncwarp();
        }

        for (int b = 0; b < p.nq; ++b) {
            const float *qs = qs_all + (size_t)b * p.wpad;
            float acc[WPT], ring[RING];
#pragma unroll
            for (int w = 0; w < WPT; ++w) acc[w] = 0.0f;
            {
                const float4 a = *reinterpret_cast<const float4 *>(yb + 0);
                const float4 c = *reinterpret_cast<const float4 *>(yb + 4);
                const float4 e = *reinterpret_cast<const float4 *>(yb + 8);
                ring[0] = a.x; ring[1] = a.y; ring[2] = a.z; ring[3] = a.w;
                ring[4] = c.x; ring[5] = c.y; ring[6] = c.z; ring[7] = c.w;
                ring[8] = e.x; ring[9] = e.y; ring[10] = e.z; ring[11] = e.w;
                ring[12] = ring[13] = ring[14] = ring[15] = 0.0f;
            }
            int j0 = 0;
#pragma unroll 1
            for (; j0 + RING <= p.W; j0 += RING) step_block<EXACT, false>(acc, ring, yb + j0, qs + j0, RING);
            if (j0 < p.W) step_block<EXACT, true>(acc, ring, yb + j0, qs + j0, p.W - j0);

            // ---- epilogue: rare candidates are appended to the per-query lists ----
            unsigned int mask = 0;
            if (EXACT) {
                const float s_thr = ld_volatile_f32(&p.st[b].s_thr);
#pragma unroll
                for (int w = 0; w < WPT; ++w)
                    if (acc[w] <= s_thr) mask |= 1u << w;
            } else {
                // LB <= thr_fast  <=>  (Y2^ - 2 D^) + (Q2 - slack - thr_fast) <= 0
                const float q2 = p.st[b].q2;
                const float thr = ld_volatile_f32(&p.st[b].thr_fast);
                const float base = (q2 - p.cw * (q2 + ptot)) - thr;
#pragma unroll
                for (int w = 0; w < WPT; ++w) {
                    const float v = fmaf(-2.0f, acc[w], y2[w]) + base;
                    if (!(v > 0.0f)) mask |= 1u << w;
                }
            }
            if (!all_valid) {
#pragma unroll
                for (int w = 0; w < WPT; ++w)
                    if (tl + w >= tk.tp_eff) mask &= ~(1u << w);
            }
            if (__any_sync(FULL, mask != 0)) {
                unsigned long long key[WPT];
                if (EXACT) {
                    const float qn = p.st[b].qnorm;
                    const unsigned long long tau = ld_volatile_u64(&p.st[b].tau_key);
#pragma unroll
                    for (int w = 0; w < WPT; ++w) {
                        key[w] = 0;
                        if (mask & (1u << w)) {
                            const float d = dist_from_s(acc[w], qn);
                            key[w] = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned long long)(flat0 + w);
                            if (key[w] > tau) mask &= ~(1u << w);
                        }
                    }
                }
                const int cnt = __popc(mask);
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += v;
                }
                const int total = __shfl_sync(FULL, incl, 31);
                if (total > 0) {
                    unsigned int base = 0;
                    if (lane == 31) base = atomicAdd(EXACT ? &p.st[b].count : &p.st[b].ccount, (unsigned int)total);
                    base = __shfl_sync(FULL, base, 31);
                    unsigned int pos = base + (unsigned int)(incl - cnt);
                    if (EXACT) {
                        unsigned long long *dst = p.keys + ((size_t)b * 2 + p.st[b].cur) * p.cap;
#pragma unroll
                        for (int w = 0; w < WPT; ++w)
                            if (mask & (1u << w)) {
                                if (pos < p.cap) dst[pos] = key[w];
                                ++pos;
                            }
                    } else {
                        unsigned int *dst = p.cand + (size_t)b * p.cap;
#pragma unroll
                        for (int w = 0; w < WPT; ++w)
                            if (mask & (1u << w)) {
                                if (pos < p.cap) dst[pos] = flat0 + w;
                                ++pos;
                            }
                    }
                }
            }
        }
        __syncwarp();  // every lane is done with buf before it is refilled
        if (!have_next) break;
        tk = tn;
        task = next;
    }
}

#include "pshadow_embed.cuh"

// ------------------------------------------------------------------------------------------
// FFT flavour: preparation kernels, query spectrum, scan
// ------------------------------------------------------------------------------------------
#include "pshadow_fftscan.cuh"
#include "pshadow_fft3.cuh"

#include "pshadow_embed_fft.cuh"

// ------------------------------------------------------------------------------------------
// exact re-rank of the filter's candidates.  Each warp takes 32 candidates at a time: their
// windows are staged into a shared-memory tile with coalesced 4-byte cp.async copies (windows
// are only 4-byte aligned), then lane c evaluates candidate c with the reference's sequential
// sub/mul/add chain (tile row stride is odd: conflict-free), applies the exact thresholds and
// appends the (distance bits, flat index) key.  grid = (blocks, nq).
// ------------------------------------------------------------------------------------------
constexpr int RR_WARPS = 4;
constexpr int RR_THREADS = RR_WARPS * 32;
constexpr int RR_JC = 128;          // reduction steps staged per pass (66 KB per CTA: three CTAs per SM, so the re-rank of a pipelined
                                    // scan gets through on the two SMs the neighbouring stream's scan leaves free)
constexpr int RR_WP = RR_JC + 1;    // tile row stride (floats)


__global__ void __launch_bounds__(RR_THREADS) rerank_kernel(const float *__restrict__ ds, long long row_stride,
                                                             unsigned int Tp, int W,
                                                             const float *__restrict__ queries, QState *st_all,
                                                             const unsigned int *__restrict__ cand_all,
                                                             unsigned long long *keys_all, unsigned int cap) {
    extern __shared__ __align__(16) float rr_smem[];
    const int wq = (W + 3) & ~3;
    float *qsh = rr_smem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *tile = rr_smem + wq + (size_t)warp * 32 * RR_WP;

    const int b = blockIdx.y;
    QState *st = st_all + b;
    const unsigned int craw = st->ccount;
    const unsigned int C = min(craw, cap);
    if (craw > cap && threadIdx.x == 0 && blockIdx.x == 0) { st->overflow = 1; st->sticky = 1; }
    const unsigned int groups = (C + 31u) / 32u;
    if (blockIdx.x * RR_WARPS >= groups) return;
    for (int j = threadIdx.x; j < W; j += RR_THREADS) qsh[j] = queries[(size_t)b * W + j];
    __syncthreads();
    const float s_thr = st->s_thr, qn = st->qnorm;
    const unsigned long long tau = st->tau_key;
    const unsigned int *cand = cand_all + (size_t)b * cap;
    unsigned long long *dst = keys_all + ((size_t)b * 2 + st->cur) * cap;

    for (unsigned int g = blockIdx.x * RR_WARPS + warp; g < groups; g += gridDim.x * RR_WARPS) {
        const unsigned int i = g * 32u + lane;
        const bool valid = i < C;
        const unsigned int flat = valid ? cand[i] : 0u;
        const unsigned int r = flat / Tp, t = flat - r * Tp;
        const long long off = (long long)r * row_stride + t;
        const unsigned int vmask = __ballot_sync(FULL, valid);
        float s = 0.0f;
        for (int j0 = 0; j0 < W; j0 += RR_JC) {
            const int jn = min(RR_JC, W - j0);
            for (int c = 0; c < 32; ++c) {
                const long long oc = __shfl_sync(FULL, off, c);
                if (vmask & (1u << c)) {
                    const float *src = ds + oc + j0;
                    const uint32_t d0 = smem_u32(tile + c * RR_WP);
                    for (int j = lane; j < jn; j += 32) cp_async_4(d0 + 4u * j, src + j);
                }
            }
            cp_async_wait_all();
            __syncwarp();
            if (valid) {
                const float *row = tile + lane * RR_WP;
                const float *qq = qsh + j0;
#pragma unroll 4
                for (int j = 0; j < jn; ++j) {
                    const float df = __fsub_rn(qq[j], row[j]);
                    s = __fadd_rn(s, __fmul_rn(df, df));
                }
            }
            __syncwarp();
        }
        bool keep = false;
        unsigned long long key = 0;
        if (valid && s <= s_thr) {
            key = ((unsigned long long)__float_as_uint(dist_from_s(s, qn)) << 32) | flat;
            keep = key <= tau;
        }
        const unsigned int bal = __ballot_sync(FULL, keep);
        if (bal) {
            unsigned int base = 0;
            if (lane == 0) base = atomicAdd(&st->count, (unsigned int)__popc(bal));
            base = __shfl_sync(FULL, base, 0);
            const unsigned int pos = base + __popc(bal & ((1u << lane) - 1u));
            if (keep && pos < cap) dst[pos] = key;
        }
    }
}

// ------------------------------------------------------------------------------------------
// bitonic sort of n (power of two) 64-bit keys by one CTA; data in shared or global memory
// ------------------------------------------------------------------------------------------
__device__ void bitonic_sort_cta(unsigned long long *a, unsigned int n) {
    for (unsigned int size = 2; size <= n; size <<= 1) {
        for (unsigned int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (unsigned int i = threadIdx.x; i < (n >> 1); i += blockDim.x) {
                unsigned int lo = 2 * i - (i & (stride - 1));
                unsigned int hi = lo + stride;
                bool up = (lo & size) == 0;
                unsigned long long x = a[lo], y = a[hi];
                if ((x > y) == up) { a[lo] = y; a[hi] = x; }
            }
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------
// select: keep the k smallest keys of a query's candidate list, publish the new thresholds and,
// after the last chunk, order the k keys and decode them into the outputs.
// Replaces torch.topk + cat + topk (path_shadowing.py:165,170-173).  One CTA per query.
//   fast path   : min/max of the distance bits -> 2048 linear bins over that range -> the bin
//                 holding the k-th key -> keys below it are kept, keys inside it (a handful) are
//                 sorted in shared memory and the smallest ones complete the k  (3 passes)
//   generic path: MSB-first radix select on the full 64-bit key (8-bit digits), used when the
//                 fast path degenerates (all distances equal, or a huge tie group at the k-th)
// ------------------------------------------------------------------------------------------
// results: separate (B,k) distances + (B,k,2) indices, or -- when out_idx is NULL -- packed
// (B,k,3) int32 records [distance bits, trajectory, offset] in out_d (the all-gather payload)
__device__ __forceinline__ void write_result(float *out_d, int *out_idx, size_t pos, unsigned int dbits, int r, int t) {
    if (out_idx != nullptr) {
        out_d[pos] = __uint_as_float(dbits);
        out_idx[pos * 2 + 0] = r;
        out_idx[pos * 2 + 1] = t;
    } else {
        int *rec = reinterpret_cast<int *>(out_d) + pos * 3;
        rec[0] = (int)dbits; rec[1] = r; rec[2] = t;
    }
}

constexpr int SEL_BINS = 2048;
constexpr int SEL_LIST = 4096;  // boundary-bin keys / fused final sort capacity (keys)
constexpr int SEL_KPT = 24;     // keys per thread cached in registers (24 k keys per query)

__device__ __forceinline__ void hist_add(unsigned int *hist, unsigned int digit, bool active) {
    // warp-aggregated shared-memory histogram increment
    unsigned int act = __ballot_sync(FULL, active);
    if (!active) return;
    unsigned int peers = __match_any_sync(act, digit);
    if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[digit], (unsigned int)__popc(peers));
}

// generic path; returns tau (exactly k keys of src are <= tau) and compacts them into dst
__device__ unsigned long long select_generic(const unsigned long long *src, unsigned long long *dst, unsigned int M,
                                             unsigned int k, unsigned int *hist, unsigned long long *s_prefix,
                                             unsigned int *s_need, unsigned int *s_done, unsigned int *s_out) {
    const int tid = threadIdx.x;
    if (tid == 0) { *s_prefix = 0; *s_need = k; *s_done = 0; *s_out = 0; }
    __syncthreads();
    unsigned long long tau = ~0ull;
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = 56 - 8 * pass;
        for (int i = tid; i < 256; i += SEL_THREADS) hist[i] = 0;
        __syncthreads();
        const unsigned long long prefix = *s_prefix;
        const unsigned int Mr = (M + 31u) & ~31u;
        for (unsigned int i = tid; i < Mr; i += SEL_THREADS) {
            bool act = i < M;
            unsigned long long key = act ? src[i] : 0ull;
            if (pass > 0) act = act && ((key >> (shift + 8)) == (prefix >> (shift + 8)));
            hist_add(hist, (unsigned int)(key >> shift) & 255u, act);
        }
        __syncthreads();
        if (tid < 32) {  // warp 0: locate the bin holding the need-th smallest key
            unsigned int need = *s_need;
            unsigned int loc[8], sum = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) { loc[i] = hist[tid * 8 + i]; sum += loc[i]; }
            unsigned int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned int v = __shfl_up_sync(FULL, incl, o);
                if (tid >= o) incl += v;
            }
            unsigned int excl = incl - sum;
            __syncwarp();  // every lane has read *s_need / *s_prefix before one lane rewrites them
            if (excl < need && need <= incl) {  // exactly one lane
                unsigned int cc = excl;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (cc < need && need <= cc + loc[i]) {
                        *s_prefix = prefix | ((unsigned long long)(tid * 8 + i) << shift);
                        *s_need = need - cc;
                        *s_done = (loc[i] == need - cc) ? 1u : 0u;
                    }
                    cc += loc[i];
                }
            }
        }
        __syncthreads();
        if (*s_done || pass == 7) {
            tau = *s_prefix | (shift ? ((1ull << shift) - 1ull) : 0ull);
            break;
        }
    }
    for (unsigned int i = tid; i < ((M + 31u) & ~31u); i += SEL_THREADS) {
        bool keep = false;
        unsigned long long key = 0;
        if (i < M) { key = src[i]; keep = key <= tau; }
        unsigned int bal = __ballot_sync(FULL, keep);
        if (bal) {
            unsigned int base = 0;
            if ((tid & 31) == 0) base = atomicAdd(s_out, (unsigned int)__popc(bal));
            base = __shfl_sync(FULL, base, 0);
            if (keep) dst[base + __popc(bal & ((1u << (tid & 31)) - 1u))] = key;
        }
    }
    __syncthreads();
    return tau;
}

__global__ void __launch_bounds__(SEL_THREADS) select_kernel(QState *st_all, unsigned long long *keys_all,
                                                              unsigned int cap, unsigned int k, int W, int final_sort,
                                                              unsigned int Tp, int row_offset, float *out_d,
                                                              int *out_idx, unsigned int *fft_hist) {
    __shared__ unsigned int hist[SEL_BINS];
    if (fft_hist != nullptr)  // fresh in-launch threshold histogram for the next FFT round
        for (int i = threadIdx.x; i < HB + HC; i += SEL_THREADS) fft_hist[(size_t)blockIdx.x * HSTRIDE + i] = 0u;
    if (fft_hist != nullptr && threadIdx.x == 0) fft_hist[(size_t)blockIdx.x * HSTRIDE + H_SLOT] = 0u;   // the next launch's pair slots
    __shared__ unsigned long long list[SEL_LIST];
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned int s_need, s_done, s_out, s_min, s_max, s_bin, s_below, s_nlist, s_fast;
    QState *st = st_all + blockIdx.x;
    const unsigned int cnt_raw = st->count;
    const unsigned int M = min(cnt_raw, cap);
    const unsigned int cur = st->cur;
    const unsigned long long *src = keys_all + ((size_t)blockIdx.x * 2 + cur) * cap;
    unsigned long long *dst = keys_all + ((size_t)blockIdx.x * 2 + (cur ^ 1)) * cap;
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) {
        if (cnt_raw > cap) { st->overflow = 1; st->sticky = 1; }
        s_min = 0xffffffffu; s_max = 0u; s_out = 0; s_nlist = 0; s_fast = 0;
    }
    const unsigned long long *kept = src;  // where the k (or M <= k) surviving keys live
    unsigned int nkept = M;
    if (M > k) {  // uniform branch
        for (int i = tid; i < SEL_BINS; i += SEL_THREADS) hist[i] = 0;
        // the first SEL_KPT*1024 keys are read from global memory ONCE (independent loads, one
        // latency) and kept in registers for all three passes; longer lists re-read the tail
        unsigned long long kreg[SEL_KPT];
#pragma unroll
        for (int r = 0; r < SEL_KPT; ++r) {
            const unsigned int i = tid + r * SEL_THREADS;
            kreg[r] = i < M ? src[i] : ~0ull;
        }
        __syncthreads();
        // pass 1: range of the distance bits
        unsigned int mn = 0xffffffffu, mx = 0u;
#pragma unroll
        for (int r = 0; r < SEL_KPT; ++r) {
            if (tid + r * SEL_THREADS < M) {
                const unsigned int db = (unsigned int)(kreg[r] >> 32);
                mn = min(mn, db); mx = max(mx, db);
            }
        }
        for (unsigned int i = tid + SEL_KPT * SEL_THREADS; i < M; i += SEL_THREADS) {
            const unsigned int db = (unsigned int)(src[i] >> 32);
            mn = min(mn, db); mx = max(mx, db);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = min(mn, __shfl_xor_sync(FULL, mn, o));
            mx = max(mx, __shfl_xor_sync(FULL, mx, o));
        }
        if (lane == 0) { atomicMin(&s_min, mn); atomicMax(&s_max, mx); }
        __syncthreads();
        const unsigned int lo = s_min, range = s_max - s_min;
        unsigned int shift = 0;
        while ((range >> shift) >= (unsigned int)SEL_BINS) ++shift;
        bool fast = range > 0;
        if (fast) {
            // pass 2: histogram
#pragma unroll
            for (int r = 0; r < SEL_KPT; ++r)
                if (tid + r * SEL_THREADS < M) atomicAdd(&hist[((unsigned int)(kreg[r] >> 32) - lo) >> shift], 1u);
            for (unsigned int i = tid + SEL_KPT * SEL_THREADS; i < M; i += SEL_THREADS)
                atomicAdd(&hist[((unsigned int)(src[i] >> 32) - lo) >> shift], 1u);
            __syncthreads();
            if (tid < 32) {  // the bin holding the k-th key
                unsigned int sum = 0;
                for (int i = 0; i < SEL_BINS / 32; ++i) sum += hist[tid * (SEL_BINS / 32) + i];
                unsigned int incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    unsigned int v = __shfl_up_sync(FULL, incl, o);
                    if (tid >= o) incl += v;
                }
                unsigned int excl = incl - sum;
                if (excl < k && k <= incl) {
                    unsigned int cc = excl;
                    for (int i = 0; i < SEL_BINS / 32; ++i) {
                        const unsigned int h = hist[tid * (SEL_BINS / 32) + i];
                        if (cc < k && k <= cc + h) { s_bin = tid * (SEL_BINS / 32) + i; s_below = cc; s_fast = h <= (unsigned int)SEL_LIST; }
                        cc += h;
                    }
                }
            }
            __syncthreads();
            fast = s_fast != 0;
        }
        unsigned long long tau;
        if (fast) {
            // pass 3: keys below the boundary bin are kept, keys inside it go to the list
            const unsigned int kb = s_bin;
            const unsigned int Mr = (M + 31u) & ~31u;
#pragma unroll
            for (int r = 0; r < SEL_KPT; ++r) {
                const unsigned int i = tid + r * SEL_THREADS;
                if (i - lane < Mr) {  // warp-uniform
                    const unsigned long long key = kreg[r];
                    const unsigned int bin = i < M ? (((unsigned int)(key >> 32) - lo) >> shift) : 0xffffffffu;
                    const bool keep = bin < kb;
                    if (bin == kb) list[atomicAdd(&s_nlist, 1u)] = key;
                    const unsigned int bal = __ballot_sync(FULL, keep);
                    if (bal) {
                        unsigned int base = 0;
                        if (lane == 0) base = atomicAdd(&s_out, (unsigned int)__popc(bal));
                        base = __shfl_sync(FULL, base, 0);
                        if (keep) dst[base + __popc(bal & ((1u << lane) - 1u))] = key;
                    }
                }
            }
            for (unsigned int i = tid + SEL_KPT * SEL_THREADS; i < Mr; i += SEL_THREADS) {
                unsigned long long key = 0;
                unsigned int bin = 0xffffffffu;
                if (i < M) { key = src[i]; bin = ((unsigned int)(key >> 32) - lo) >> shift; }
                const bool keep = bin < kb;
                if (bin == kb) list[atomicAdd(&s_nlist, 1u)] = key;
                const unsigned int bal = __ballot_sync(FULL, keep);
                if (bal) {
                    unsigned int base = 0;
                    if (lane == 0) base = atomicAdd(&s_out, (unsigned int)__popc(bal));
                    base = __shfl_sync(FULL, base, 0);
                    if (keep) dst[base + __popc(bal & ((1u << lane) - 1u))] = key;
                }
            }
            __syncthreads();
            const unsigned int nl = s_nlist, need = k - s_below;  // 1 <= need <= nl
            unsigned int np2 = 1; while (np2 < nl) np2 <<= 1;
            for (unsigned int i = nl + tid; i < np2; i += SEL_THREADS) list[i] = ~0ull;
            bitonic_sort_cta(list, np2);
            tau = list[need - 1];
            for (unsigned int i = tid; i < need; i += SEL_THREADS) dst[s_below + i] = list[i];
            __syncthreads();
            if (tid == 0) s_out = k;
        } else {
            tau = select_generic(src, dst, M, k, hist, &s_prefix, &s_need, &s_done, &s_out);
        }
        kept = dst;
        nkept = k;
        if (tid < 32) {
            // s_thr: the largest float s with dist_from_s(s) <= tau's distance (monotone map)
            const float qn = st->qnorm;
            const float taud = __uint_as_float((unsigned int)(tau >> 32));
            float s_thr;
            if (!(taud < __int_as_float(0x7f800000)) || !(qn > 0.0f)) {
                s_thr = __int_as_float(0x7f800000);
            } else {
                const float est = __fmul_rn(__fmul_rn(taud, qn), __fmul_rn(taud, qn));
                unsigned int eb = __float_as_uint(est);
                // probe est-16 .. est+15 ulps in parallel, fall back to bisection outside that band
                unsigned int lo_b = eb > 16u ? eb - 16u : 0u;
                unsigned int cand = min(lo_b + (unsigned int)tid, 0x7f800000u);
                bool ok = dist_from_s(__uint_as_float(cand), qn) <= taud;
                unsigned int okm = __ballot_sync(FULL, ok);
                if (okm != 0u && okm != FULL) {
                    int hi = 31 - __clz(okm);  // monotone: ok lanes form a prefix
                    s_thr = __uint_as_float(min(lo_b + (unsigned int)hi, 0x7f800000u));
                } else {
                    unsigned int l2 = 0u, h2 = 0x7f800000u;  // invariant: f(l2) ok (s=0 -> d=0), answer in [l2,h2]
                    while (l2 < h2) {
                        unsigned int mid = l2 + (h2 - l2 + 1u) / 2u;
                        if (dist_from_s(__uint_as_float(mid), qn) <= taud) l2 = mid; else h2 = mid - 1u;
                    }
                    s_thr = __uint_as_float(l2);
                }
            }
            if (tid == 0) {
                st->tau_key = tau;
                st->s_thr = s_thr;
                // exact s >= S_true (1 - gamma_{W+2}) - W 2^-126  =>  S_true <= thr_fast (rounded up)
                const double widen = 1.0 + 2.0 * (double)(W + 8) * 5.9604644775390625e-8;
                st->thr_fast = (s_thr < __int_as_float(0x7f800000))
                                   ? __double2float_ru((double)s_thr * widen + 1e-30)
                                   : s_thr;
                st->count = k;
                st->ccount = 0;
                st->cur = cur ^ 1u;
            }
        }
    } else {
        if (tid == 0) { st->count = M; st->ccount = 0; }
    }
    if (!final_sort) return;
    // fused finalisation (k <= SEL_LIST): order the surviving keys, decode (r, t)
    __syncthreads();
    unsigned int np2 = 1; while (np2 < k) np2 <<= 1;
    for (unsigned int i = tid; i < np2; i += SEL_THREADS) list[i] = i < nkept ? kept[i] : ~0ull;
    bitonic_sort_cta(list, np2);
    for (unsigned int i = tid; i < k; i += SEL_THREADS) {
        const unsigned long long key = list[i];
        const unsigned int flat = (unsigned int)key;
        // packed output of a NOSYNC scan: an overflowed query poisons its first record so that every
        // rank sees it after the all-gather (merge_kernel raises the flag)
        const bool poison = out_idx == nullptr && i == 0 && st->overflow != 0;
        write_result(out_d, out_idx, (size_t)blockIdx.x * k + i, poison ? 0xffffffffu : (unsigned int)(key >> 32),
                     (int)(flat / Tp) + row_offset, (int)(flat % Tp));
    }
}

// final ordering of a query's k keys and decoding into (distance, [trajectory, offset])
__global__ void __launch_bounds__(SEL_THREADS) finalize_kernel(const QState *st_all, unsigned long long *keys_all,
                                                                unsigned int cap, unsigned int k, unsigned int npow2,
                                                                int use_smem, unsigned int Tp, int row_offset,
                                                                float *out_d, int *out_idx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const QState *st = st_all + blockIdx.x;
    unsigned long long *src = keys_all + ((size_t)blockIdx.x * 2 + st->cur) * cap;
    unsigned long long *other = keys_all + ((size_t)blockIdx.x * 2 + (st->cur ^ 1u)) * cap;
    const unsigned int M = min(st->count, cap);
    unsigned long long *a = use_smem ? reinterpret_cast<unsigned long long *>(smem_raw) : other;
    for (unsigned int i = threadIdx.x; i < npow2; i += blockDim.x) a[i] = i < M ? src[i] : ~0ull;
    bitonic_sort_cta(a, npow2);
    for (unsigned int i = threadIdx.x; i < k; i += blockDim.x) {
        unsigned long long key = a[i];
        unsigned int flat = (unsigned int)key;
        // packed output of a NOSYNC scan: an overflowed query poisons its first record so that every
        // rank sees it after the all-gather (merge_kernel raises the flag)
        const bool poison = out_idx == nullptr && i == 0 && st->overflow != 0;
        write_result(out_d, out_idx, (size_t)blockIdx.x * k + i, poison ? 0xffffffffu : (unsigned int)(key >> 32),
                     (int)(flat / Tp) + row_offset, (int)(flat % Tp));
    }
}

// ------------------------------------------------------------------------------------------
// merge of G sorted shard results (multi-GPU): path_shadowing.py:170-173 across ranks
// keys are rebuilt as (distance bits, global r*Tp+t) -- 96 bits do not fit one word, so the
// sort key is (dbits << 32 | slot) with slot = g*k+i and ties in distance are re-ordered by the
// global flat index in a final pass over equal-distance runs.
// ------------------------------------------------------------------------------------------
// Record j of shard g, query b: distance at d_parts[rec * dstride], indices at i_parts[rec * istride + {0,1}],
// rec = (g*B + b)*k + j  (separate arrays: dstride 1, istride 2; packed [dbits, r, t] records: 3, 3).
__device__ __forceinline__ void merge_body(const float *d_parts, const int *i_parts, int dstride, int istride, int G,
                                           int B, unsigned int k, unsigned long long Tp, unsigned int npow2,
                                           unsigned long long *scratch, int use_smem, float *out_d, int *out_idx,
                                           int *flag, unsigned char *smem_raw) {
    const int b = blockIdx.x;
    const unsigned int n = (unsigned int)G * k;
    if (use_smem == 2) {
        // merge by rank: every shard's list is already ascending in (distance bits, global flat index),
        // so the output position of element j of list g is j + sum over the other lists of the number
        // of their elements ordered before it (one binary search each) -- no sort, no tie fix-up.
        // Elements equal in distance AND flat index (the +inf padding of short shards) order by g.
        unsigned long long *fl = reinterpret_cast<unsigned long long *>(smem_raw);
        unsigned int *db = reinterpret_cast<unsigned int *>(fl + n);
        for (unsigned int i = threadIdx.x; i < n; i += blockDim.x) {
            const unsigned int g = i / k, j = i - g * k;
            const size_t rec = ((size_t)g * B + b) * k + j;
            const unsigned int d = __float_as_uint(__ldcg(d_parts + rec * dstride));
            if (flag != nullptr && d == 0xffffffffu) atomicOr(flag, 1);  // a shard overflowed
            const int *ic = i_parts + rec * istride;
            db[i] = d;
            fl[i] = (unsigned long long)(unsigned int)__ldcg(ic) * Tp + (unsigned long long)(unsigned int)__ldcg(ic + 1);
        }
        __syncthreads();
        for (unsigned int i = threadIdx.x; i < n; i += blockDim.x) {
            const unsigned int g = i / k, j = i - g * k;
            const unsigned int dk = db[i];
            const unsigned long long fk = fl[i];
            unsigned int rank = j;
            for (unsigned int g2 = 0; g2 < (unsigned int)G && rank < k; ++g2) {
                if (g2 == g) continue;
                const unsigned int base = g2 * k;
                unsigned int lo = 0, hi = k;
                while (lo < hi) {
                    const unsigned int mid = (lo + hi) >> 1;
                    const unsigned int d2 = db[base + mid];
                    bool before = d2 < dk;
                    if (d2 == dk) {
                        const unsigned long long f2 = fl[base + mid];
                        before = f2 < fk || (f2 == fk && g2 < g);
                    }
                    if (before) lo = mid + 1; else hi = mid;
                }
                rank += lo;
            }
            if (rank < k) {
                const int *ic = i_parts + (((size_t)g * B + b) * k + j) * istride;
                out_d[(size_t)b * k + rank] = __uint_as_float(dk);
                out_idx[((size_t)b * k + rank) * 2 + 0] = __ldcg(ic);
                out_idx[((size_t)b * k + rank) * 2 + 1] = __ldcg(ic + 1);
            }
        }
        return;
    }
    unsigned long long *a = use_smem ? reinterpret_cast<unsigned long long *>(smem_raw)
                                     : scratch + (size_t)b * npow2;
    for (unsigned int i = threadIdx.x; i < npow2; i += blockDim.x) {
        unsigned long long key = ~0ull;
        if (i < n) {
            unsigned int g = i / k, j = i - g * k;
            // L2 loads: in the peer-memory exchange these records were written by other GPUs
            float d = __ldcg(d_parts + (((size_t)g * B + b) * k + j) * dstride);
            if (flag != nullptr && __float_as_uint(d) == 0xffffffffu) atomicOr(flag, 1);  // a shard overflowed
            key = ((unsigned long long)__float_as_uint(d) << 32) | i;
        }
        a[i] = key;
    }
    bitonic_sort_cta(a, npow2);
    // fix the order inside runs of equal distance: rank by global flat index (runs are tiny)
    for (unsigned int i = threadIdx.x; i < k; i += blockDim.x) {
        unsigned long long key = a[i];
        unsigned int db = (unsigned int)(key >> 32);
        unsigned int lo = i, hi = i;
        while (lo > 0 && (unsigned int)(a[lo - 1] >> 32) == db) --lo;
        while (hi + 1 < n && (unsigned int)(a[hi + 1] >> 32) == db) ++hi;
        unsigned int slot = (unsigned int)key;
        if (lo != hi) {
            // rank of this element's flat index among the run -> the element that belongs at i
            unsigned int want = i - lo;
            for (unsigned int c = lo; c <= hi; ++c) {
                unsigned int sc = (unsigned int)a[c];
                unsigned int gc = sc / k, jc = sc - gc * k;
                const int *ic = i_parts + (((size_t)gc * B + b) * k + jc) * istride;
                unsigned long long fc = (unsigned long long)__ldcg(ic) * Tp + (unsigned long long)__ldcg(ic + 1);
                unsigned int rank = 0;
                for (unsigned int e = lo; e <= hi; ++e) {
                    unsigned int se = (unsigned int)a[e];
                    unsigned int ge = se / k, je = se - ge * k;
                    const int *ie = i_parts + (((size_t)ge * B + b) * k + je) * istride;
                    unsigned long long fe = (unsigned long long)__ldcg(ie) * Tp + (unsigned long long)__ldcg(ie + 1);
                    rank += (fe < fc) ? 1u : 0u;
                }
                if (rank == want) { slot = sc; break; }
            }
        }
        unsigned int g = slot / k, j = slot - g * k;
        const int *src = i_parts + (((size_t)g * B + b) * k + j) * istride;
        out_d[(size_t)b * k + i] = __uint_as_float(db);
        out_idx[((size_t)b * k + i) * 2 + 0] = __ldcg(src);
        out_idx[((size_t)b * k + i) * 2 + 1] = __ldcg(src + 1);
    }
}

__global__ void __launch_bounds__(SEL_THREADS) merge_kernel(const float *d_parts, const int *i_parts, int dstride,
                                                             int istride, int G, int B,
                                                             unsigned int k, unsigned long long Tp, unsigned int npow2,
                                                             unsigned long long *scratch, int use_smem,
                                                             float *out_d, int *out_idx, int *flag) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    merge_body(d_parts, i_parts, dstride, istride, G, B, k, Tp, npow2, scratch, use_smem, out_d, out_idx, flag, smem_raw);
}

// ------------------------------------------------------------------------------------------
// all-gather over NVLink peer memory fused with the merge (multi-GPU, one process per GPU).
// Every rank owns an exchange buffer that all other ranks have mapped (CUDA IPC):
//     [epoch % 8][ records (G, B, k, 3) int32 | flags (G, B) uint32 ]
// CTA b of rank r stores query b's k packed records into slot r of EVERY rank's buffer (plain
// stores over NVLink), fences (system scope), then raises flag (r, b) on every rank with the
// step's epoch; it then waits until the G flags of query b in its OWN buffer carry the epoch
// and merges the G*k records exactly like merge_kernel.  One launch replaces ncclAllGather +
// merge_kernel.  Eight buffers rotate with the epoch (see XCHG_PARITIES).
// A peer that never arrives (crashed rank) raises bit 1 of `flag` after `timeout_ns` instead of
// hanging the GPU.
// ------------------------------------------------------------------------------------------
constexpr int XCHG_MAX_PEERS = 16;
// Buffers rotate over 4 epochs.  One stream of fused send+merge launches needs 2 (a rank's step e+2 follows
// its step e+1, which needed every peer's send e+1, which follows that peer's merge e).  Two streams that
// alternate steps need 2 per stream, and so does the split form (send e+1 enqueued before merge e): 8 buffers
// cover pipelines of up to four streams.
constexpr int XCHG_PARITIES = 8;
struct XchgParams {
    int *rec[XCHG_MAX_PEERS];            // this parity's record area on every rank
    unsigned int *flags[XCHG_MAX_PEERS]; // this parity's flag area on every rank
    const int *local_rec;                // (B, k, 3) this rank's records
    int G, rank;
    unsigned int epoch;
    unsigned long long timeout_ns;
};

__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(SEL_THREADS) xchg_merge_kernel(const XchgParams x, int B, unsigned int k,
                                                                  unsigned long long Tp, unsigned int npow2,
                                                                  unsigned long long *scratch, int use_smem,
                                                                  float *out_d, int *out_idx, int *flag, int phases) {
    // phases: bit 0 = send (stores + flags), bit 1 = wait + merge.  Split, the two halves of a step
    // can be enqueued around the NEXT step's scan, so a rank never idles waiting for a slower peer.
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x, tid = threadIdx.x;
    const unsigned int n3 = k * 3u;
    if (phases & 1) {
        const int *src = x.local_rec + (size_t)b * n3;
        for (int g = 0; g < x.G; ++g) {
            int *dst = x.rec[g] + ((size_t)x.rank * B + b) * n3;
            for (unsigned int i = tid; i < n3; i += SEL_THREADS) dst[i] = src[i];
        }
        __syncthreads();
        if (tid < x.G) {
            __threadfence_system();   // the CTA's record stores (ordered before by the barrier) precede the flag
            st_release_sys(x.flags[tid] + (size_t)x.rank * B + b, x.epoch);
        }
    }
    if (!(phases & 2)) return;
    if (tid < x.G) {
        const unsigned int *mine = x.flags[x.rank] + (size_t)tid * B + b;
        const unsigned long long t0 = globaltimer_ns();
        while (ld_acquire_sys(mine) != x.epoch) {
            if (globaltimer_ns() - t0 > x.timeout_ns) { if (flag != nullptr) atomicOr(flag, 2); break; }
            __nanosleep(100);
        }
    }
    __syncthreads();
    const int *all = x.rec[x.rank];
    merge_body(reinterpret_cast<const float *>(all), all + 1, 3, 3, x.G, B, k, Tp, npow2, scratch, use_smem, out_d,
               out_idx, flag, smem_raw);
}

// ------------------------------------------------------------------------------------------
// The same exchange without fences or flags ("LL": every 4-byte datum travels in ONE 8-byte store
// together with the step's epoch, so a word validates itself -- the receiver polls each word until
// its upper half carries the epoch; 8-byte stores are single-copy atomic, also over NVLink).  The
// system-scope fence + flag round trip of xchg_merge_kernel is what its 35-45 us were spent on.
// Merge by rank straight out of the polled words: (distance bits, flat index) go to shared
// memory, the output (r, t) is recovered from the flat index.  Needs G*k*12 bytes of shared memory
// (else xchg_merge_kernel runs).  The record area of the exchange buffer is sized for this form.
// ------------------------------------------------------------------------------------------
// grid = (G, B): CTA (g, b) stores query b's records into rank g's buffer, then -- like its G - 1 siblings,
// each on its own SM -- polls ALL G record lists of query b out of this rank's buffer (the loads of a batch
// of records are issued together and validated afterwards: one L2 round trip per batch, not per word) and
// places the records of list g only (merge by rank: position = own index + the number of records of the
// other lists ordered before it).  Round 1 ran ONE CTA per query: G sequential remote-store loops, word-by-
// word polling and the whole merge on one SM (0.076 ms at G = 8).
constexpr int LL_BATCH = 4;   // records polled per thread and round trip (12 words in flight)
constexpr int LL_THREADS = 512; // (46 registers: two CTAs per SM)

__global__ void __launch_bounds__(LL_THREADS) xchg_ll_kernel(const XchgParams x, int B, unsigned int k,
                                                               unsigned long long Tp, float *out_d, int *out_idx,
                                                               int *flag, int phases) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int g = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const unsigned int n3 = k * 3u;
    const unsigned long long tag = (unsigned long long)x.epoch << 32;
    if (phases & 1) {
        const int *src = x.local_rec + (size_t)b * n3;
        volatile unsigned long long *dst =
            reinterpret_cast<volatile unsigned long long *>(x.rec[g]) + ((size_t)x.rank * B + b) * n3;
        for (unsigned int i = tid; i < n3; i += LL_THREADS) dst[i] = tag | (unsigned int)src[i];
    }
    if (!(phases & 2)) return;
    const unsigned int n = (unsigned int)x.G * k;
    unsigned long long *fl = reinterpret_cast<unsigned long long *>(smem_raw);
    unsigned int *db = reinterpret_cast<unsigned int *>(fl + n);
    const volatile unsigned long long *mine = reinterpret_cast<const volatile unsigned long long *>(x.rec[x.rank]);
    const unsigned long long t0 = globaltimer_ns();
    bool timed_out = false;
    for (unsigned int i0 = tid; i0 < n; i0 += LL_THREADS * LL_BATCH) {
        unsigned long long word[LL_BATCH][3];
        const volatile unsigned long long *wp[LL_BATCH];
#pragma unroll
        for (int u = 0; u < LL_BATCH; ++u) {
            const unsigned int i = i0 + u * LL_THREADS;
            const unsigned int gg = i < n ? i / k : 0u, j = i < n ? i - gg * k : 0u;
            wp[u] = mine + (((size_t)gg * B + b) * k + j) * 3;
        }
#pragma unroll
        for (int u = 0; u < LL_BATCH; ++u)
#pragma unroll
            for (int c = 0; c < 3; ++c) word[u][c] = wp[u][c];
#pragma unroll
        for (int u = 0; u < LL_BATCH; ++u) {
            const unsigned int i = i0 + u * LL_THREADS;
            if (i >= n) continue;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                while ((word[u][c] >> 32) != x.epoch && !timed_out) {
                    if (globaltimer_ns() - t0 > x.timeout_ns) { timed_out = true; break; }
                    __nanosleep(64);
                    word[u][c] = wp[u][c];
                }
            }
            const unsigned int v0 = (unsigned int)word[u][0];
            if (flag != nullptr && v0 == 0xffffffffu && g == 0) atomicOr(flag, 1);   // a shard overflowed
            db[i] = v0;
            fl[i] = (unsigned long long)(unsigned int)word[u][1] * Tp + (unsigned long long)(unsigned int)word[u][2];
        }
    }
    if (timed_out && flag != nullptr) atomicOr(flag, 2);
    __syncthreads();
    for (unsigned int j = tid; j < k; j += LL_THREADS) {   // the records of list g
        const unsigned int i = (unsigned int)g * k + j;
        const unsigned int dk = db[i];
        const unsigned long long fk = fl[i];
        unsigned int rank = j;
        for (unsigned int g2 = 0; g2 < (unsigned int)x.G && rank < k; ++g2) {
            if (g2 == (unsigned int)g) continue;
            const unsigned int base = g2 * k;
            unsigned int lo = 0, hi = k;
            while (lo < hi) {
                const unsigned int mid = (lo + hi) >> 1;
                const unsigned int d2 = db[base + mid];
                bool before = d2 < dk;
                if (d2 == dk) {
                    const unsigned long long f2 = fl[base + mid];
                    before = f2 < fk || (f2 == fk && g2 < (unsigned int)g);
                }
                if (before) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        if (rank < k) {
            out_d[(size_t)b * k + rank] = __uint_as_float(dk);
            out_idx[((size_t)b * k + rank) * 2 + 0] = (int)(fk / Tp);
            out_idx[((size_t)b * k + rank) * 2 + 1] = (int)(fk % Tp);
        }
    }
}

// ------------------------------------------------------------------------------------------
// gather: paths[i, :] = dataset[r, t : t+L]   (path_shadowing.py:210-216), one warp per path
// ------------------------------------------------------------------------------------------
__global__ void gather_kernel(const float *__restrict__ ds, long long R, long long row_stride,
                              const int *__restrict__ idx, long long n, int row_offset, int L,
                              float *__restrict__ out) {
    long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const int lane = threadIdx.x & 31;
    long long r = (long long)idx[2 * i] - row_offset;
    const int t = idx[2 * i + 1];
    float *o = out + i * L;
    if (r < 0 || r >= R) {
        for (int j = lane; j < L; j += 32) o[j] = 0.0f;
        return;
    }
    const float *src = ds + r * row_stride + t;
    for (int j = lane; j < L; j += 32) o[j] = __ldg(src + j);
}

// ------------------------------------------------------------------------------------------
// realised variance + weighted aggregation: one CTA per query
//   x_i(path) = 252 * mean(path_out[:T_i]^2) (sqrt if vol)             statistics.py:5-16
//   w(path) ~ exp(-d^2 / (2 eta^2)) (softmax) or 1/k (uniform)          scatspectra (unpinned)
//   mean_i = sum w x_i ; std_i = sqrt(sum w x_i^2 - mean_i^2)           path_shadowing.py:248-252
// accumulated in fp64 (k*H values per query: negligible work), rounded once to fp32.
// ------------------------------------------------------------------------------------------
constexpr int AGG_THREADS = 256;
constexpr int AGG_MAX_T = 16;

__device__ __forceinline__ double block_sum(double v, double *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int i = 0; i < AGG_THREADS / 32; ++i) s += red[i];
    return s;
}

__global__ void __launch_bounds__(AGG_THREADS) rv_aggregate_kernel(const float *__restrict__ paths,
                                                                   const float *__restrict__ dist, long long k, int L,
                                                                   int H, const int *__restrict__ Ts, int nT, float eta,
                                                                   int proba, int vol, float *out_mean, float *out_std) {
    __shared__ double red[AGG_THREADS / 32];
    __shared__ int sT[AGG_MAX_T];
    const int b = blockIdx.x;
    if (threadIdx.x < nT) sT[threadIdx.x] = Ts[threadIdx.x];
    __syncthreads();
    const float *dq = dist + (size_t)b * k;
    // minimum squared distance: the shift that keeps exp() in range (cancels in the ratio)
    double dmin = 1e300;
    for (long long i = threadIdx.x; i < k; i += AGG_THREADS) {
        double d = (double)dq[i];
        dmin = fmin(dmin, d * d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dmin = fmin(dmin, __shfl_xor_sync(FULL, dmin, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dmin;
    __syncthreads();
    dmin = red[0];
    for (int i = 1; i < AGG_THREADS / 32; ++i) dmin = fmin(dmin, red[i]);
    const double inv2e2 = proba == 1 ? 1.0 / (2.0 * (double)eta * (double)eta) : 0.0;

    double sw = 0.0, swx[AGG_MAX_T], swxx[AGG_MAX_T];
#pragma unroll
    for (int i = 0; i < AGG_MAX_T; ++i) { swx[i] = 0.0; swxx[i] = 0.0; }
    for (long long i = threadIdx.x; i < k; i += AGG_THREADS) {
        const float *o = paths + ((size_t)b * k + i) * L + (L - H);
        double d = (double)dq[i];
        double w = proba == 1 ? exp(-(d * d - dmin) * inv2e2) : 1.0;
        sw += w;
        double run = 0.0;
        int j = 0;
#pragma unroll
        for (int ti = 0; ti < AGG_MAX_T; ++ti) {
            if (ti < nT) {
                // maturities are sorted ascending by the host wrapper: extend the running sum
                // numpy's x2[..., :T] clips at the out-context length (statistics.py:13)
                const int Te = sT[ti] < H ? sT[ti] : H;
                for (; j < Te; ++j) { double v = (double)o[j]; run += v * v; }
                double x = run / (double)Te * 252.0;
                if (vol) x = sqrt(x);
                swx[ti] += w * x;
                swxx[ti] += w * x * x;
            }
        }
    }
    sw = block_sum(sw, red);
    for (int ti = 0; ti < nT; ++ti) {
        double a = block_sum(swx[ti < AGG_MAX_T ? ti : 0], red) / sw;
        double c = block_sum(swxx[ti < AGG_MAX_T ? ti : 0], red) / sw;
        if (threadIdx.x == 0) {
            double var = c - a * a;
            out_mean[(size_t)b * nT + ti] = (float)a;
            out_std[(size_t)b * nT + ti] = (float)sqrt(var > 0.0 ? var : 0.0);
        }
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

long long gcd_ll(long long a, long long b) { while (b) { long long t = a % b; a = b; b = t; } return a; }

// stride of the row permutation: ~0.618 R, coprime with R, so every prefix of slots is a
// spread-out sample of the ensemble (the early chunks seed the threshold)
long long perm_stride(long long R) {
    if (R <= 2) return 1;
    long long p = (long long)(0.6180339887498949 * (double)R);
    if (p < 1) p = 1;
    while (gcd_ll(p, R) != 1) ++p;
    return p % R ? p % R : 1;
}

struct Plan {
    long long Tp;
    unsigned long long N;
    unsigned int cap;
    long long n0;      // rows of the seeding chunk
    int growth;
    size_t off_state, off_keys, off_cand, off_qspec, off_hist, off_qmaxp, total;
};

constexpr int SEED_FACTOR = 16;  // seeding chunk holds ~16 k windows
constexpr int GROWTH = 20;       // each later chunk multiplies the scanned prefix by at most this
constexpr int CAP_SLACK = 4;     // candidate buffer: 4x the expected appends of a chunk

bool make_plan(long long R, long long T, int B, int W, int H, long long k, Plan &pl) {
    if (R <= 0 || T <= 0 || B <= 0 || W <= 0 || H < 0 || k <= 0) return false;
    pl.Tp = T - W - H + 1;
    if (pl.Tp <= 0) return false;
    pl.N = (unsigned long long)R * (unsigned long long)pl.Tp;
    long long n0 = (SEED_FACTOR * k + pl.Tp - 1) / pl.Tp;
    if (n0 < 1) n0 = 1;
    if (n0 > R) n0 = R;
    pl.n0 = n0;
    pl.growth = GROWTH;
    unsigned long long cap = (unsigned long long)n0 * pl.Tp + (unsigned long long)k
                             + (unsigned long long)CAP_SLACK * GROWTH * (unsigned long long)k;
    if (cap < 2ull * (unsigned long long)k) cap = 2ull * k;
    // power-of-two room for the global-memory sort of very large k
    unsigned long long p2 = 1; while (p2 < (unsigned long long)k) p2 <<= 1;
    if (k > SORT_SMEM_MAX && cap < p2) cap = p2;
    if (cap > 0xfffffff0ull) return false;
    pl.cap = (unsigned int)cap;
    pl.off_state = 0;
    pl.off_keys = align_up((size_t)B * sizeof(QState), 256);
    pl.off_cand = pl.off_keys + (size_t)B * 2 * (size_t)pl.cap * sizeof(unsigned long long);
    pl.off_qspec = pl.off_cand + align_up((size_t)B * (size_t)pl.cap * sizeof(unsigned int), 256);
    pl.off_hist = pl.off_qspec + (size_t)B * fftx::N * sizeof(float2);  // query spectra (fft flavour)
    pl.off_qmaxp = pl.off_hist + (size_t)B * HSTRIDE * sizeof(unsigned int);  // threshold histograms + published thresholds
    pl.total = pl.off_qmaxp + align_up((size_t)B * QMAXP * sizeof(float), 256);  // partial maxima of the query spectra
    return true;
}

int g_sm_count = 0;
int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}

// pinned host staging for the end-of-call status read (per thread, grown on demand, never freed)
QState *host_stage(int n) {
    static thread_local QState *buf = nullptr;
    static thread_local int cap = 0;
    if (n > cap) {
        if (buf != nullptr) cudaFreeHost(buf);
        buf = nullptr; cap = 0;
        const int want = n < QG_MAX ? QG_MAX : n;
        if (cudaHostAlloc(reinterpret_cast<void **>(&buf), sizeof(QState) * (size_t)want, cudaHostAllocDefault) != cudaSuccess) {
            buf = nullptr;
            cudaGetLastError();
            return nullptr;
        }
        cap = want;
    }
    return buf;
}

// merge by rank needs 12 bytes of shared memory per record (G*k records per query)
constexpr unsigned long long MERGE_RANK_SMEM_MAX = 200 * 1024;
bool merge_sort_forced() {   // PSH_MERGE_SORT=1: always the bitonic-sort merge (tests / A-B measurements)
    const char *e = getenv("PSH_MERGE_SORT");
    return e != nullptr && e[0] == '1';
}

// A/B switch for measurements: PSH_SEEDLESS=0 keeps the exact-seeded chunk schedule everywhere
bool seedless_enabled() {
    const char *e = getenv("PSH_SEEDLESS");  // read per call: tests toggle it
    return !(e != nullptr && e[0] == '0');
}

// optional per-kernel timing (bench.py's roofline leg): CUDA events on the caller's stream
struct ProfRec { cudaEvent_t a, b; int kind; };
thread_local bool g_prof_on = false;          // per calling thread, like host_stage(): the ABI keeps no
thread_local std::vector<ProfRec> g_prof;     // state shared between threads except the launch counter
struct ProfScope {
    cudaStream_t s; int kind; cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(cudaStream_t s_, int kind_) : s(s_), kind(kind_) {
        if (g_prof_on) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, s); }
    }
    ~ProfScope() {
        if (a) { cudaEventRecord(b, s); g_prof.push_back({a, b, kind}); }
    }
};

#define PSH_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)
#define PSH_LAUNCHED() do { g_launches.fetch_add(1, std::memory_order_relaxed); cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return (int)e_; } while (0)

}  // namespace

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

int psh_version(void) { return PSH_VERSION; }

const char *psh_error_string(int code) {
    switch (code) {
        case PSH_OK: return "ok";
        case PSH_E_ARG: return "invalid argument (null pointer, non-positive size, or W+H > T)";
        case PSH_E_K: return "k exceeds the number of windows";
        case PSH_E_WORKSPACE: return "workspace too small or misaligned";
        case PSH_E_TOO_LARGE: return "more than 2^32-1 windows in one call: shard the rows and merge";
        case PSH_E_UNSUPPORTED: return "context length exceeds the shared-memory budget of the scan";
        case PSH_E_OVERFLOW: return "a candidate buffer overflowed in a PSH_FLAG_NOSYNC scan: repeat the call without the flag";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown pshadow error";
}

uint64_t psh_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

void psh_profile_begin(void) {
    for (auto &r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof.clear();
    g_prof_on = true;
}

int psh_profile_end(double *ms_by_kind, uint64_t *launches_by_kind, int nkinds) {
    g_prof_on = false;
    for (int i = 0; i < nkinds; ++i) { ms_by_kind[i] = 0.0; launches_by_kind[i] = 0; }
    int rc = PSH_OK;
    for (auto &r : g_prof) {
        float ms = 0.f;
        cudaError_t e = cudaEventSynchronize(r.b);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, r.a, r.b);
        if (e != cudaSuccess) rc = (int)e;
        if (r.kind >= 0 && r.kind < nkinds) { ms_by_kind[r.kind] += ms; launches_by_kind[r.kind] += 1; }
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    g_prof.clear();
    return rc;
}

size_t psh_scan_workspace_bytes(int64_t R, int64_t T, int B, int W, int H, int64_t k) {
    Plan pl;
    if (!make_plan(R, T, B, W, H, k, pl)) return 0;
    return pl.total;
}

size_t psh_fft_aux_bytes(int64_t R, int64_t T, int W, int H) {
    FftAux a;
    if (!fft_aux_layout(R, T, W, H, nullptr, a)) return 0;
    return a.total;
}

// Transform length every prepared aux buffer was laid out with (keyed by its device address): a scan uses the
// length of ITS buffer even if the PSH_FFT_N knob changed between psh_fft_prepare and the scan.
static std::mutex g_aux_mu;
static std::unordered_map<const void *, int> g_aux_nfft;
static void aux_nfft_record(const void *d_aux, int nfft) {
    std::lock_guard<std::mutex> lock(g_aux_mu);
    if (g_aux_nfft.size() > 65536) g_aux_nfft.clear();   // (stale addresses of freed buffers: bounded)
    g_aux_nfft[d_aux] = nfft;
}
static int aux_nfft_lookup(const void *d_aux) {
    std::lock_guard<std::mutex> lock(g_aux_mu);
    auto it = g_aux_nfft.find(d_aux);
    return it == g_aux_nfft.end() ? 0 : it->second;
}

// spectra + pair statistics, then the energy table: window energies (Identity) or, with a run table, the
// embedded energies E2 = sum_n e_n(t)^2
static int fft_prepare_impl(const float *d_dataset, int64_t R, int64_t T, int64_t row_stride, int W, int H,
                            void *d_aux, size_t aux_bytes, const EmbRun *d_runs, int nruns, cudaStream_t stream) {
    if (!d_dataset || !d_aux || row_stride < T) return PSH_E_ARG;
    FftAux a;
    if (!fft_aux_layout(R, T, W, H, static_cast<unsigned char *>(d_aux), a)) return W > fftx::N / 2 ? PSH_E_UNSUPPORTED : PSH_E_ARG;
    if (aux_bytes < a.total || (reinterpret_cast<uintptr_t>(d_aux) & 255u)) return PSH_E_WORKSPACE;
    aux_nfft_record(d_aux, a.nfft);
    const size_t smem = (size_t)nruns * sizeof(EmbRun);
    if (smem > 12 * 1024) return PSH_E_UNSUPPORTED;   // next to the 32 KiB fp64 prefix array
    // twiddle tables: exp(+2 pi i m / 4096) in fp64, rounded once for the fp32 copy
    static std::vector<double2> h64;
    static std::vector<float2> h32;
    static std::vector<float4> h2;   // 1024-point flavour: {w^(lane 2j), w^(lane (2j+1))} at [j * 32 + lane]
    static std::mutex tw_mutex;
    {
        std::lock_guard<std::mutex> lock(tw_mutex);
        if (h64.empty()) {
            h64.resize(fftx::N); h32.resize(fftx::N); h2.resize(512);
            for (int m = 0; m < fftx::N; ++m) {
                const double ang = 6.283185307179586476925286766559 * (double)m / (double)fftx::N;
                h64[m].x = cos(ang); h64[m].y = sin(ang);
                h32[m].x = (float)h64[m].x; h32[m].y = (float)h64[m].y;
            }
            for (int j = 0; j < 16; ++j)
                for (int lane = 0; lane < 32; ++lane) {
                    const int m0 = (lane * 2 * j) & 1023, m1 = (lane * (2 * j + 1)) & 1023;   // powers of exp(2 pi i / 1024)
                    h2[j * 32 + lane] = make_float4(h32[4 * m0].x, h32[4 * m0].y, h32[4 * m1].x, h32[4 * m1].y);
                }
        }
    }
    PSH_CUDA(cudaMemcpyAsync(a.tw64, h64.data(), sizeof(double2) * fftx::N, cudaMemcpyHostToDevice, stream));
    PSH_CUDA(cudaMemcpyAsync(a.tw32, h32.data(), sizeof(float2) * fftx::N, cudaMemcpyHostToDevice, stream));
    PSH_CUDA(cudaMemcpyAsync(a.tw2, h2.data(), sizeof(float4) * 512, cudaMemcpyHostToDevice, stream));
    const int Tp = (int)(T - W - H + 1);
    if (a.nfft == fx3::N) {
        fft3_prep_spectra_kernel<<<(unsigned int)((a.npairs + 3) / 4), 128, 0, stream>>>(d_dataset, (int)T, row_stride, a);
        PSH_LAUNCHED();
        if (d_runs != nullptr)
            fft_prep_energy_kernel<true, 1024><<<(unsigned int)a.npairs, fftx::THREADS, smem, stream>>>(
                d_dataset, (int)T, row_stride, W, Tp, a, d_runs, nruns);
        else
            fft_prep_energy_kernel<false, 1024><<<(unsigned int)a.npairs, fftx::THREADS, 0, stream>>>(
                d_dataset, (int)T, row_stride, W, Tp, a, nullptr, 0);
        PSH_LAUNCHED();
        return PSH_OK;
    }
    fft_prep_spectra_kernel<<<(unsigned int)a.npairs, fftx::THREADS, 0, stream>>>(d_dataset, (int)T, row_stride, a);
    PSH_LAUNCHED();
    if (d_runs != nullptr)
        fft_prep_energy_kernel<true, 4096><<<(unsigned int)a.npairs, fftx::THREADS, smem, stream>>>(
            d_dataset, (int)T, row_stride, W, Tp, a, d_runs, nruns);
    else
        fft_prep_energy_kernel<false, 4096><<<(unsigned int)a.npairs, fftx::THREADS, 0, stream>>>(
            d_dataset, (int)T, row_stride, W, Tp, a, nullptr, 0);
    PSH_LAUNCHED();
    return PSH_OK;
}

int psh_fft_prepare(const float *d_dataset, int64_t R, int64_t T, int64_t row_stride, int W, int H,
                    void *d_aux, size_t aux_bytes, void *stream_) {
    return fft_prepare_impl(d_dataset, R, T, row_stride, W, H, d_aux, aux_bytes, nullptr, 0, (cudaStream_t)stream_);
}

int psh_debug_fft4096(const void *d_in, void *d_out, int n, int dir, const void *d_aux, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!d_in || !d_out || !d_aux || n <= 0) return PSH_E_ARG;
    fft_debug_kernel<<<n, fftx::THREADS, 0, stream>>>(static_cast<const float2 *>(d_in), static_cast<float2 *>(d_out),
                                                      static_cast<const float2 *>(d_aux), dir);
    PSH_LAUNCHED();
    return PSH_OK;
}

int psh_debug_fft1024(const void *d_in, void *d_out, int n, int dir, const void *d_aux, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!d_in || !d_out || !d_aux || n <= 0 || (dir != 1 && dir != -1)) return PSH_E_ARG;
    // the twiddle table of the 1024-point flavour sits behind the two 4096-entry tables of a prepared aux buffer
    const float4 *tw2 = reinterpret_cast<const float4 *>(static_cast<const unsigned char *>(d_aux)
                                                         + (sizeof(float2) + sizeof(double2)) * fftx::N);
    fft3_debug_kernel<<<(n + 3) / 4, 128, 0, stream>>>(static_cast<const float2 *>(d_in), static_cast<float2 *>(d_out), n, tw2, dir);
    PSH_LAUNCHED();
    return PSH_OK;
}

static __global__ void clear_sticky_kernel(QState *st, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st[i].sticky = 0u;
}

// final ordering of the k keys when it was not fused into the last select (k > SEL_LIST)
static int launch_finalize(const Plan &pl, QState *st, unsigned long long *keys, int nq, long long k, int row_offset,
                           float *d_out_dist, int *d_out_idx, cudaStream_t stream) {
    unsigned int npow2 = 1; while (npow2 < (unsigned int)k) npow2 <<= 1;
    int use_smem = npow2 <= SORT_SMEM_MAX ? 1 : 0;
    size_t fsmem = use_smem ? (size_t)npow2 * sizeof(unsigned long long) : 0;
    {
        ProfScope ps(stream, 1);
        finalize_kernel<<<nq, SEL_THREADS, fsmem, stream>>>(st, keys, pl.cap, (unsigned int)k, npow2, use_smem,
                                                            (unsigned int)pl.Tp, row_offset, d_out_dist, d_out_idx);
    }
    PSH_LAUNCHED();
    return PSH_OK;
}

// ---- one-time per-device setup: dynamic shared-memory limits, resident CTAs of the fft scan ----
constexpr size_t SMEM_BIG = 200 * 1024;          // budget of the kernels whose shared memory grows with W / k
constexpr size_t SMEM_FFT = sizeof(__half2) * 2 * fx2::N + sizeof(float2) * fx2::EX_FLOAT2 + sizeof(float4) + 16;
typedef void (*FftScanFn)(const FftScanParams);
struct FftVariant { FftScanFn fn; int ctas_per_sm; };
constexpr int FFT_VARIANTS = 4;                  // [query spectrum in registers][emb]
constexpr size_t SMEM_FFT3_SINGLE = fx3::TW_BYTES + fx3::Q_BYTES + (size_t)fx3::WARPS_SINGLE * fx3::WARP_BYTES_ALIAS;
constexpr size_t SMEM_FFT3_GROUP = fx3::TW_BYTES + (size_t)fx3::WARPS_GROUP * fx3::WARP_BYTES_SEP;
struct DevSetup { bool done = false; FftVariant fft[FFT_VARIANTS]; };
static DevSetup g_dev[64];

static FftScanFn fft_variant_fn(int sq, int em) {
    if (sq && em) return fft_scan_kernel<true, true>;
    if (sq) return fft_scan_kernel<true, false>;
    if (em) return fft_scan_kernel<false, true>;
    return fft_scan_kernel<false, false>;
}

// [embedded][one query][energy groups per piece: 7 / run-time]
static FftScanFn fft3_variant_fn(int em, int single, int ncy) {
    if (ncy == 7) {
        if (single) return em ? fft_scan_warp_kernel<true, true, 7> : fft_scan_warp_kernel<false, true, 7>;
        return em ? fft_scan_warp_kernel<true, false, 7> : fft_scan_warp_kernel<false, false, 7>;
    }
    if (single) return em ? fft_scan_warp_kernel<true, true, 0> : fft_scan_warp_kernel<false, true, 0>;
    return em ? fft_scan_warp_kernel<true, false, 0> : fft_scan_warp_kernel<false, false, 0>;
}

#define big_smem(kernel, bytes) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))

static int device_setup(DevSetup **out) {
    int dev = 0;
    PSH_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return PSH_E_ARG;
    DevSetup &d = g_dev[dev];
    *out = &d;
    if (d.done) return PSH_OK;
    PSH_CUDA(big_smem(scan_kernel<true>, SMEM_BIG));
    PSH_CUDA(big_smem(scan_kernel<false>, SMEM_BIG));
    PSH_CUDA(big_smem(emb_scan_kernel<1>, SMEM_BIG));
    PSH_CUDA(big_smem(emb_scan_kernel<EMB_QG>, SMEM_BIG));
    PSH_CUDA(big_smem(rerank_kernel, SMEM_BIG));
    PSH_CUDA(big_smem(emb_rerank_kernel, SMEM_BIG));
    PSH_CUDA(big_smem(finalize_kernel, SMEM_BIG));
    PSH_CUDA(big_smem(merge_kernel, SMEM_BIG));
    PSH_CUDA(big_smem(xchg_merge_kernel, SMEM_BIG));
    PSH_CUDA(big_smem(xchg_ll_kernel, SMEM_BIG));
    // CUDA loads a kernel's module lazily at its first launch, which synchronises the whole context: a
    // first launch issued while an exchange kernel of another stream waits for a peer would stall behind it
    // (and two ranks doing so for each other's sake would only be released by the exchange's timeout).
    // Touch every kernel now, before any exchange kernel can be resident.
    {
        cudaFuncAttributes fa;
        PSH_CUDA(cudaFuncGetAttributes(&fa, qprep_kernel));
        PSH_CUDA(cudaFuncGetAttributes(&fa, qfft_kernel));
        PSH_CUDA(cudaFuncGetAttributes(&fa, select_kernel));
        PSH_CUDA(cudaFuncGetAttributes(&fa, gather_kernel));
        PSH_CUDA(cudaFuncGetAttributes(&fa, rv_aggregate_kernel));
        PSH_CUDA(cudaFuncGetAttributes(&fa, clear_sticky_kernel));
        PSH_CUDA(cudaFuncGetAttributes(&fa, fft_prep_spectra_kernel));
        PSH_CUDA(cudaFuncGetAttributes(&fa, fft_prep_energy_kernel<false, 4096>));
        PSH_CUDA(cudaFuncGetAttributes(&fa, fft_prep_energy_kernel<true, 4096>));
        PSH_CUDA(cudaFuncGetAttributes(&fa, fft_prep_energy_kernel<false, 1024>));
        PSH_CUDA(cudaFuncGetAttributes(&fa, fft_prep_energy_kernel<true, 1024>));
        PSH_CUDA(cudaFuncGetAttributes(&fa, fft3_prep_spectra_kernel));
    }
    for (int i = 0; i < 8; ++i)
        PSH_CUDA(big_smem(fft3_variant_fn(i & 1, (i >> 1) & 1, (i >> 2) & 1 ? 7 : 0), ((i >> 1) & 1) ? SMEM_FFT3_SINGLE : SMEM_FFT3_GROUP));
    for (int i = 0; i < FFT_VARIANTS; ++i) {
        FftScanFn fn = fft_variant_fn((i >> 1) & 1, i & 1);
        PSH_CUDA(big_smem(fn, SMEM_FFT));
        int nb = 0;
        PSH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, fx2::THREADS, SMEM_FFT));
        d.fft[i].fn = fn;
        d.fft[i].ctas_per_sm = nb > 0 ? nb : 1;
    }
    d.done = true;
    return PSH_OK;
}

static int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return (e != nullptr && e[0] != 0) ? atoi(e) : dflt;
}

// shared memory of the direct-evaluation kernels for a group of nq queries
struct DirectSmem { int epl, pfx_floats, buf_floats, wpad; size_t exact, filter; };
static DirectSmem direct_smem(int W, int qlen, int nq) {
    DirectSmem m;
    const int need = SEG + W - 1;
    int epl4 = ((need + 31) / 32 + 3) / 4;
    if ((epl4 & 1) == 0) ++epl4;  // odd multiple of 4 floats per lane: conflict-free LDS.128/STS.128
    m.epl = 4 * epl4;
    m.pfx_floats = 4 + 32 * m.epl + 4;
    m.buf_floats = (int)align_up((size_t)need + RING + 4, 4);
    if (m.buf_floats < 32 * m.epl + 4) m.buf_floats = 32 * m.epl + 4;
    m.wpad = (int)align_up((size_t)qlen + 4, 4);
    m.exact = ((size_t)nq * m.wpad + (size_t)SCAN_WARPS * 2 * m.buf_floats) * sizeof(float)
              + (size_t)SCAN_WARPS * 2 * sizeof(unsigned long long);
    m.filter = m.exact + (size_t)SCAN_WARPS * m.pfx_floats * sizeof(float);
    return m;
}
static size_t emb_smem(int W, int d, int nruns, int nq) {
    const DirectSmem m = direct_smem(W, d, nq);
    const int ps_n = (int)align_up((size_t)SEG + W + 1, 2);
    return (size_t)nq * m.wpad * sizeof(float) + (size_t)nruns * sizeof(EmbRun) + (size_t)SCAN_WARPS * ps_n * sizeof(float2)
           + (size_t)SCAN_WARPS * 2 * m.buf_floats * sizeof(float) + (size_t)SCAN_WARPS * 2 * sizeof(unsigned long long);
}
// queries per scan group: as many (<= QG_MAX) as the shared memory of the kernels this call may launch
// holds; 0: not even one query fits (context too long for the staging buffers)
static int query_group_size(int W, int mode, bool has_aux, const EmbParams *emb) {
    for (int nq = QG_MAX; nq >= 1; --nq) {
        if (emb) {
            if (emb_smem(W, emb->d, emb->nruns, nq) <= SMEM_BIG) return nq;
            continue;
        }
        const DirectSmem m = direct_smem(W, W, nq);
        const bool filter_used = mode == PSH_MODE_FILTER || (mode == PSH_MODE_FFT && !has_aux);
        if ((filter_used ? m.filter : m.exact) <= SMEM_BIG) return nq;
    }
    return 0;
}

static int run_scan_group(const float *d_dataset, long long R, long long T, long long row_stride,
                          const float *d_q, int nq, int W, int H, long long k, int row_offset,
                          const Plan &pl, QState *st, unsigned long long *keys, unsigned int *cand, float2 *qspec,
                          unsigned int *fhist, float *qmaxp, const FftAux *aux, int mode, bool safe,
                          float *d_out_dist, int *d_out_idx, cudaStream_t stream, const EmbParams *emb = nullptr,
                          int spare_sms = 0) {
    (void)H;
    DevSetup *dv = nullptr;
    { int rc_ = device_setup(&dv); if (rc_ != PSH_OK) return rc_; }
    const bool use_fft = (mode == PSH_MODE_FFT) && !safe && aux != nullptr;
    const bool filter = (mode == PSH_MODE_FILTER || (mode == PSH_MODE_FFT && !use_fft)) && !safe;
    if (use_fft && emb && emb->g == nullptr) return PSH_E_ARG;
    const int qlen = emb ? emb->d : W;   // embedded scan: d_q holds the EMBEDDED queries (nq, d)
    if (use_fft) {
        // query state + spectrum of the vector the trajectories are correlated with (the context itself,
        // or g = K^T ex for an embedded scan) + a clean threshold histogram: ONE launch
        qfft_kernel<<<dim3(aux->nfft / QFFT_K, nq), QFFT_THREADS, (size_t)W * sizeof(double), stream>>>(
            d_q, qlen, emb ? emb->g : d_q, W, aux->tw64, qspec, st, fhist, qmaxp, aux->nfft);
    } else {
        qprep_kernel<<<(nq + 3) / 4, 128, 0, stream>>>(d_q, qlen, nq, st);
    }
    PSH_LAUNCHED();

    ScanParams p;
    p.ds = d_dataset; p.row_stride = row_stride; p.T = (int)T; p.Tp = (int)pl.Tp; p.W = W;
    // tasks per row; in the fft flavour per VIRTUAL row (a piece of `span` windows of a long trajectory)
    p.nsegv = use_fft ? aux->nsegv : 1;
    p.hop = use_fft ? aux->hop : 0;
    p.span = use_fft ? aux->span : (int)pl.Tp;
    p.VR = use_fft ? aux->VR : R;
    p.nseg = (int)((p.span + SEG - 1) / SEG);
    p.R = R; p.perm = perm_stride(R);
    p.queries = d_q; p.nq = nq; p.st = st; p.keys = keys; p.cand = cand; p.cap = pl.cap;
    p.bulk_ok = ((reinterpret_cast<uintptr_t>(d_dataset) & 15u) == 0 && (row_stride & 3) == 0) ? 1 : 0;
    const DirectSmem dm = direct_smem(W, qlen, nq);
    p.epl = dm.epl; p.pfx_floats = dm.pfx_floats; p.buf_floats = dm.buf_floats; p.wpad = dm.wpad;
    p.cw = (float)(W + 256) * 5.9604644775390625e-8f;
    EmbParams ep;
    size_t smem_emb = 0;
    if (emb) {
        ep = *emb;
        ep.ps_n = (int)align_up((size_t)SEG + W + 1, 2);
        smem_emb = emb_smem(W, emb->d, emb->nruns, nq);
        if (smem_emb > SMEM_BIG) return PSH_E_UNSUPPORTED;
    }
    const size_t smem_exact = dm.exact, smem_filter = dm.filter;
    // (scan_entry sized the query group for the kernels this call can launch)
    if (!emb && (filter ? smem_filter : smem_exact) > SMEM_BIG) return PSH_E_UNSUPPORTED;
    const size_t smem_rr = ((size_t)((W + 3) & ~3) + (size_t)RR_WARPS * 32 * RR_WP) * sizeof(float);
    if (smem_rr > SMEM_BIG) return PSH_E_UNSUPPORTED;
    auto ctas_per_sm = [](size_t smem, int lim) {
        int c = (int)((224 * 1024) / (smem + 1024));
        return c > lim ? lim : (c < 1 ? 1 : c);
    };
    {
        unsigned int sft = 0;
        while ((1u << sft) < (unsigned int)p.nseg) ++sft;
        p.nseg_s = sft;
        p.nseg_m = (unsigned int)((((1ull << 32) * ((1ull << sft) - (unsigned long long)p.nseg)) / (unsigned long long)p.nseg) + 1ull);
    }
    p.inv_R = 1.0 / (double)R;
    p.pair_mode = use_fft ? 1 : 0;
    p.npairs = use_fft ? aux->npairs : (R + 1) / 2;
    p.inv_np = 1.0 / (double)p.npairs;
    if (use_fft) p.perm = perm_stride(p.npairs);
    FftScanParams fp;
    FftVariant fv = {nullptr, 1};
    if (use_fft) {
        fp.Z = aux->Z; fp.Y2 = aux->Y2; fp.pinfo = aux->pinfo; fp.tw = aux->tw32; fp.Qc = qspec; fp.qmaxp = qmaxp;
        fp.Tp = (int)pl.Tp; fp.nq = nq;
        fp.npairs = (int)p.npairs; fp.VR = aux->VR; fp.nsegv = aux->nsegv; fp.hop = aux->hop;
        fp.perm = p.perm; fp.inv_np = p.inv_np;
        fp.st = st; fp.cand = cand; fp.cap = pl.cap;
        fp.cf_u = 512.0f * 5.9604644775390625e-8f;
        fp.hist = fhist; fp.k = (unsigned int)k;
        fp.seed = 0; fp.seed_need = 1;
        fp.tw2 = aux->tw2; fp.ncy = aux->ncy; fp.nqmax = aux->nfft / QFFT_K;
        fp.stagger_ns = (unsigned int)env_int("PSH_FFT_STAGGER_NS", 0);
        fp.dbg = nullptr;
        { const char *e_ = getenv("PSH_FFT_DBG"); if (e_ != nullptr && e_[0] != 0) fp.dbg = reinterpret_cast<unsigned long long *>(strtoull(e_, nullptr, 0)); }
        fp.refresh_mask = (unsigned int)env_int("PSH_FFT_REFRESH", aux->nfft == fx3::N ? 31 : 7);
        {
            const double w1 = 1.0 + 2.0 * (double)(qlen + 8) * 5.9604644775390625e-8;
            const double we = emb ? 1.0 + 1.0 / 512.0 : 1.0;   // the exact embedded evaluation vs the true S
            fp.widen2 = (float)(w1 * w1 * we * we * (1.0 + 1e-6));
            fp.thr_widen = (float)we;
            // Identity: 12u (Q2 + ynorm^2) covers the roundings of Q2, Y2 and of the combination (in the kernel).
            // Embedded: |2 D| <= Q2 + E2_t (Cauchy-Schwarz in embedded space), so every rounding is
            // <= a few u (Q2 + E2_t): 20u Q2 in the slack, 16u E2_t taken out of the stored energies
            // (fft_prep_energy_kernel<true>) and given back twice in UB; 2u ||g|| ynorm for the rounding of g
            fp.slack_coef = (emb ? 20.0f : 12.0f) * 5.9604644775390625e-8f;
            fp.g_coef = emb ? 2.0f * 5.9604644775390625e-8f : 0.0f;
            // UB - LB: the fp16 floor of the staged energies (2^-10) + the embedded scan's 2 x 16u
            // (1024-point flavour: bf16 energies, floor of 2^-7)
            fp.ub_y_coef = (aux->nfft == fx3::N ? 7.8125e-3f : 9.765625e-4f) * 1.01f + (emb ? 2.0f * 16.0f * 5.9604644775390625e-8f * 1.001f : 0.0f);
        }
        // one query: its spectrum streamed from L1/L2 like a group's (measured 5 % faster: no spills at 128
        // registers) or held in registers (PSH_FFT_QREG=1)
        const int qreg = nq == 1 && env_int("PSH_FFT_QREG", 0) != 0;
        fv = dv->fft[(qreg << 1) | (emb != nullptr ? 1 : 0)];
    }
    // exact re-rank of the fft / fma filter's survivors (embedded scans: emb_rerank_kernel)
    const size_t smem_er = emb ? (size_t)((emb->d + 3) & ~3) * sizeof(float) * (1 + ERR_WARPS) + (size_t)emb->nruns * sizeof(EmbRun)
                                     + (size_t)ERR_WARPS * (W + 2) * sizeof(float2)
                               : 0;
    if (smem_er > SMEM_BIG) return PSH_E_UNSUPPORTED;
    auto launch_rerank = [&]() -> int {
        ProfScope ps(stream, 1);
        if (emb) {
            unsigned int rb = (pl.cap + ERR_WARPS - 1) / ERR_WARPS;
            const unsigned int rb_max = (unsigned int)sm_count() * 8u;
            if (rb > rb_max) rb = rb_max;
            emb_rerank_kernel<<<dim3(rb, nq), ERR_WARPS * 32, smem_er, stream>>>(
                d_dataset, row_stride, (unsigned int)pl.Tp, W, emb->d, d_q, ep.runs, ep.nruns, st, cand, keys, pl.cap);
        } else {
            unsigned int rb = (pl.cap + RR_THREADS - 1) / RR_THREADS;
            const unsigned int rb_max = (unsigned int)sm_count() * 3u;
            if (rb > rb_max) rb = rb_max;
            rerank_kernel<<<dim3(rb, nq), RR_THREADS, smem_rr, stream>>>(
                d_dataset, row_stride, (unsigned int)pl.Tp, W, d_q, st, cand, keys, pl.cap);
        }
        g_launches.fetch_add(1, std::memory_order_relaxed);
        return (int)cudaGetLastError();
    };
    // 1024-point flavour: `units` are warps (one transform each), a CTA of 16 (one query) or 12 (a group) per SM
    const bool warp_fft = use_fft && aux->nfft == fx3::N;
    const int fft3_warps = nq == 1 ? fx3::WARPS_SINGLE : fx3::WARPS_GROUP;
    auto launch_fft = [&](unsigned int units) {
        ProfScope ps(stream, 0);
        if (warp_fft) {
            const unsigned int grid = (units + fft3_warps - 1) / fft3_warps;
            const size_t smem = nq == 1 ? SMEM_FFT3_SINGLE : SMEM_FFT3_GROUP;
            fft3_variant_fn(emb != nullptr, nq == 1, fp.ncy)<<<grid, fft3_warps * 32, smem, stream>>>(fp);
        } else {
            fv.fn<<<units, fx2::THREADS, SMEM_FFT, stream>>>(fp);
        }
    };
    // (a pipelined scan leaves `spare_sms` SMs to the other streams' small kernels)
    const int fft_sms = sm_count() - spare_sms > 0 ? sm_count() - spare_sms : 1;
    const long long fft_grid_max = warp_fft ? (long long)fft_sms * fft3_warps : (long long)fft_sms * fv.ctas_per_sm;
    long long fft_unit_threads = fx2::THREADS;   // seed entries per unit
    fp.seed_group = 1;
    const bool fuse_final = k <= SEL_LIST;                  // last select also sorts and decodes

    // ---- seedless schedule: ONE launch over all pairs that seeds its own threshold from every CTA's
    // first pair (fft_scan_kernel, p.seed) -> exact re-rank of the survivors -> select.  No exact seed
    // round, no intermediate select and no exact threshold exist before the final select: the re-rank
    // keeps every survivor (s_thr = +inf, tau = ~0 from the query preparation) and the select picks the
    // k best keys.  4 launches per query.
    if (use_fft && seedless_enabled()) {
        long long ctas = p.npairs < fft_grid_max ? p.npairs : fft_grid_max;
        if (warp_fft) {
            // entries per warp: as few as give 4 k entries over the launch (every entry costs two atomics on
            // a handful of addresses; the threshold's quality depends on the windows seeded, not on the entries)
            int e = env_int("PSH_FFT_SEED_ENTRIES", 0);
            if (e != 1 && e != 2 && e != 4 && e != 8 && e != 16 && e != 32) {
                e = 1;
                while (e < 32 && ctas * e < 4 * k) e *= 2;
            }
            fft_unit_threads = e;
            fp.seed_group = 32 / e;
        }
        // (k-th smallest of ctas*256 per-thread minima: with k <= half of them the hidden second-smallest
        // values of a thread cost a few per cent of threshold quality, no more)
        if (ctas >= 1 && ctas * fft_unit_threads >= 2 * k) {
            long long need = ctas / 4, kq = (2 * k + fft_unit_threads - 1) / fft_unit_threads;
            if (need < kq) need = kq;
            if (need > ctas) need = ctas;
            if (need < 1) need = 1;
            { const int e_ = env_int("PSH_FFT_SEEDNEED", 0); if (e_ > 0 && e_ <= ctas) need = e_; }
            fp.seed = 1; fp.seed_need = (unsigned int)need;
            fp.i0 = 0; fp.i1 = (int)p.npairs;
            launch_fft((unsigned int)ctas);
            PSH_LAUNCHED();
            { int rc_ = launch_rerank(); if (rc_ != 0) return rc_; }
            {
                ProfScope ps(stream, 1);
                select_kernel<<<nq, SEL_THREADS, 0, stream>>>(st, keys, pl.cap, (unsigned int)k, qlen, fuse_final ? 1 : 0,
                                                              (unsigned int)pl.Tp, row_offset, d_out_dist, d_out_idx, nullptr);
            }
            PSH_LAUNCHED();
            if (fuse_final) return PSH_OK;
            return launch_finalize(pl, st, keys, nq, k, row_offset, d_out_dist, d_out_idx, stream);
        }
    }

    // chunk schedule over permuted slots (rows, or row pairs in the fft flavour): seed chunk
    // (always exact), then geometric growth; in safe mode every chunk fits the candidate buffer
    // even if all of its windows are appended
    const long long unit = use_fft ? 2 : 1;                 // (virtual) rows per slot
    const long long nslots = use_fft ? p.npairs : R;
    const long long win_per_slot = unit * (long long)p.span; // windows a slot can contribute at most
    long long done = 0;
    long long safe_slots = ((long long)pl.cap - k) / win_per_slot;
    if (safe_slots < 1) safe_slots = 1;
    long long seed_slots = (SEED_FACTOR * k + win_per_slot - 1) / win_per_slot;
    if (seed_slots < 1) seed_slots = 1;
    if (seed_slots > nslots) seed_slots = nslots;
    // evenly geometric rounds after the seed: ratio = (nslots/seed)^(1/rounds) <= growth
    double ratio = (double)pl.growth;
    if (nslots > seed_slots) {
        const double span = (double)nslots / (double)seed_slots;
        const int rounds = (int)ceil(log(span) / log((double)pl.growth) - 1e-9);
        ratio = pow(span, 1.0 / (double)(rounds > 0 ? rounds : 1)) * (1.0 + 1e-9);
    }
    while (done < nslots) {
        long long next;
        if (safe) next = done + safe_slots;
        else if (use_fft && done > 0) next = done == seed_slots ? (long long)ceil((double)done * pl.growth) : nslots;
        else next = done == 0 ? seed_slots : (long long)ceil((double)done * ratio);
        if (next <= done) next = done + 1;
        if (next > nslots) next = nslots;
        const bool first = done == 0;
        // small rounds are cheaper on the exact kernel (no re-rank, fills the GPU with fewer rows)
        const bool exact_round = first || safe || (!use_fft && !filter) || (use_fft && (next - done) < 64);
        if (use_fft && !exact_round) {
            fp.i0 = (int)done; fp.i1 = (int)next;
            long long ctas = next - done;
            if (ctas > fft_grid_max) ctas = fft_grid_max;
            launch_fft((unsigned int)ctas);
            PSH_LAUNCHED();
        } else {
            p.i0 = done * unit; p.i1 = next * unit;
            const bool use_filter = filter && !exact_round;
            long long ntasks = (p.i1 - p.i0) * p.nseg;  // < 2^32: nseg <= Tp and R*Tp < 2^32
            p.ntasks = (unsigned int)ntasks;
            long long ctas = (ntasks + SCAN_WARPS - 1) / SCAN_WARPS;
            long long max_ctas = (long long)sm_count() * (use_filter ? ctas_per_sm(smem_filter, 3) : ctas_per_sm(smem_exact, 2));
            if (ctas > max_ctas) ctas = max_ctas;
            if (emb) {
                long long mc = (long long)sm_count() * ctas_per_sm(smem_emb, 2);
                ctas = (ntasks + SCAN_WARPS - 1) / SCAN_WARPS;
                if (ctas > mc) ctas = mc;
                ProfScope ps(stream, 0);
                if (nq == 1) emb_scan_kernel<1><<<(unsigned int)ctas, SCAN_THREADS, smem_emb, stream>>>(p, ep);
                else emb_scan_kernel<EMB_QG><<<(unsigned int)ctas, SCAN_THREADS, smem_emb, stream>>>(p, ep);
            } else if (use_filter) {
                ProfScope ps(stream, 0);
                scan_kernel<false><<<(unsigned int)ctas, SCAN_THREADS, smem_filter, stream>>>(p);
            } else {
                ProfScope ps(stream, 0);
                scan_kernel<true><<<(unsigned int)ctas, SCAN_THREADS, smem_exact, stream>>>(p);
            }
            PSH_LAUNCHED();
        }
        if (!exact_round) { int rc_ = launch_rerank(); if (rc_ != 0) return rc_; }
        {
            ProfScope ps(stream, 1);
            const int fin = (fuse_final && next == nslots) ? 1 : 0;
            select_kernel<<<nq, SEL_THREADS, 0, stream>>>(st, keys, pl.cap, (unsigned int)k, qlen, fin,
                                                          (unsigned int)pl.Tp, row_offset, d_out_dist, d_out_idx,
                                                          use_fft ? fhist : nullptr);
        }
        PSH_LAUNCHED();
        done = next;
    }
    if (fuse_final) return PSH_OK;
    return launch_finalize(pl, st, keys, nq, k, row_offset, d_out_dist, d_out_idx, stream);
}

static int scan_entry(const float *d_dataset, int64_t R, int64_t T, int64_t row_stride,
                      const float *d_queries, int B, int W, int H, int64_t k,
                      int32_t row_offset, int mode,
                      float *d_out_dist, int32_t *d_out_idx,
                      void *d_ws, size_t ws_bytes, const void *d_aux, size_t aux_bytes, void *stream_,
                      const EmbParams *emb) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const int qstride = emb ? emb->d : W;   // floats per query in d_queries
    const bool nosync = (mode & PSH_FLAG_NOSYNC) != 0;
    const int spare_arg = (mode >> 12) & 0x3f;
    const int spare_sms = (mode & PSH_FLAG_SHARE_SMS) ? env_int("PSH_SPARE_SMS", spare_arg > 0 ? spare_arg : 6) : 0;
    mode &= ~(PSH_FLAG_NOSYNC | PSH_FLAG_SHARE_SMS | (0x3f << 12));
    if (mode != PSH_MODE_EXACT && mode != PSH_MODE_FILTER && mode != PSH_MODE_FFT) return PSH_E_ARG;
    if (!d_dataset || !d_queries || !d_out_dist || !d_ws) return PSH_E_ARG;  // d_out_idx NULL: packed records
    if (row_stride < T) return PSH_E_ARG;
    Plan pl;
    if (!make_plan(R, T, B, W, H, k, pl)) {
        if (R > 0 && T > 0 && B > 0 && W > 0 && H >= 0 && k > 0 && T - W - H + 1 > 0) return PSH_E_TOO_LARGE;
        return PSH_E_ARG;
    }
    if ((unsigned long long)k > pl.N) return PSH_E_K;
    if (pl.N >= 0xffffffffull) return PSH_E_TOO_LARGE;
    if (ws_bytes < pl.total || (reinterpret_cast<uintptr_t>(d_ws) & 255u)) return PSH_E_WORKSPACE;

    unsigned char *ws = static_cast<unsigned char *>(d_ws);
    QState *st = reinterpret_cast<QState *>(ws + pl.off_state);
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(ws + pl.off_keys);
    unsigned int *cand = reinterpret_cast<unsigned int *>(ws + pl.off_cand);
    float2 *qspec = reinterpret_cast<float2 *>(ws + pl.off_qspec);
    unsigned int *fhist_all = reinterpret_cast<unsigned int *>(ws + pl.off_hist);
    float *qmaxp_all = reinterpret_cast<float *>(ws + pl.off_qmaxp);
    FftAux aux;
    const FftAux *auxp = nullptr;
    if (mode == PSH_MODE_FFT && d_aux != nullptr) {
        if (!fft_aux_layout(R, T, W, H, const_cast<unsigned char *>(static_cast<const unsigned char *>(d_aux)), aux,
                            aux_nfft_lookup(d_aux)) ||
            aux_bytes < aux.total || (reinterpret_cast<uintptr_t>(d_aux) & 255u))
            return PSH_E_WORKSPACE;
        auxp = &aux;
    }
    // queries per group: what the shared memory of the direct-evaluation kernels holds at this context
    // length (long contexts: fewer queries per pass instead of an error)
    const int QG = query_group_size(W, mode, auxp != nullptr, emb);
    if (QG == 0) return PSH_E_UNSUPPORTED;

    // per query group: the embedded scan's cross-term vectors advance with the group
    auto group_emb = [&](int g0, EmbParams &eg) -> const EmbParams * {
        if (!emb) return nullptr;
        eg = *emb;
        if (eg.g) eg.g += (size_t)g0 * W;
        return &eg;
    };
    auto run_group = [&](int g0, int nq, bool safe) -> int {
        EmbParams eg;
        return run_scan_group(d_dataset, R, T, row_stride, d_queries + (size_t)g0 * qstride, nq, W, H, k, row_offset,
                              pl, st + g0, keys + (size_t)g0 * 2 * pl.cap, cand + (size_t)g0 * pl.cap,
                              qspec + (size_t)g0 * fftx::N, fhist_all + (size_t)g0 * HSTRIDE, qmaxp_all + (size_t)g0 * QMAXP,
                              auxp, mode, safe, d_out_dist + (size_t)g0 * k * (d_out_idx ? 1 : 3),
                              d_out_idx ? d_out_idx + (size_t)g0 * k * 2 : nullptr, stream, group_emb(g0, eg),
                              safe ? 0 : spare_sms);
    };
    for (int g0 = 0; g0 < B; g0 += QG) {
        int rc = run_group(g0, B - g0 < QG ? B - g0 : QG, false);
        if (rc != PSH_OK) return rc;
    }
    if (nosync) return PSH_OK;  // the caller checks psh_scan_overflowed() before trusting the results
    // one synchronisation: did any candidate buffer overflow (adversarially ordered data)?
    QState *hst = host_stage(B);
    if (hst == nullptr) return (int)cudaErrorMemoryAllocation;
    PSH_CUDA(cudaMemcpyAsync(hst, st, sizeof(QState) * B, cudaMemcpyDeviceToHost, stream));
    PSH_CUDA(cudaStreamSynchronize(stream));
    bool redone = false;
    for (int g0 = 0; g0 < B; g0 += QG) {
        int nq = B - g0 < QG ? B - g0 : QG;
        bool ovf = false;
        for (int i = 0; i < nq; ++i) ovf = ovf || hst[g0 + i].overflow != 0;
        if (ovf) {
            int rc = run_group(g0, nq, true);
            if (rc != PSH_OK) return rc;
            redone = true;
        }
    }
    if (redone) {  // this call has dealt with its own overflow: nothing is left pending for psh_scan_overflowed
        clear_sticky_kernel<<<(B + 127) / 128, 128, 0, stream>>>(st, B);
        PSH_LAUNCHED();
        PSH_CUDA(cudaStreamSynchronize(stream));
    }
    return PSH_OK;
}

int psh_scan_topk_f32(const float *d_dataset, int64_t R, int64_t T, int64_t row_stride,
                      const float *d_queries, int B, int W, int H, int64_t k,
                      int32_t row_offset, int mode,
                      float *d_out_dist, int32_t *d_out_idx,
                      void *d_ws, size_t ws_bytes, const void *d_aux, size_t aux_bytes, void *stream_) {
    return scan_entry(d_dataset, R, T, row_stride, d_queries, B, W, H, k, row_offset, mode, d_out_dist, d_out_idx,
                      d_ws, ws_bytes, d_aux, aux_bytes, stream_, nullptr);
}

int psh_scan_topk_embed_f32(const float *d_dataset, int64_t R, int64_t T, int64_t row_stride,
                            const float *d_qemb, int B, int d, int W, int H, int64_t k,
                            int32_t row_offset, int flags, const void *d_runs, int nruns,
                            const float *d_g, const void *d_aux, size_t aux_bytes,
                            float *d_out_dist, int32_t *d_out_idx, void *d_ws, size_t ws_bytes, void *stream_) {
    if (!d_runs || nruns <= 0 || d <= 0 || (flags & ~(PSH_FLAG_NOSYNC | PSH_FLAG_SHARE_SMS | (0x3f << 12))) != 0) return PSH_E_ARG;
    if (d_aux != nullptr && d_g == nullptr) return PSH_E_ARG;
    EmbParams ep;
    ep.runs = static_cast<const EmbRun *>(d_runs);
    ep.nruns = nruns;
    ep.d = d;
    ep.ps_n = 0;
    ep.g = d_aux ? d_g : nullptr;
    return scan_entry(d_dataset, R, T, row_stride, d_qemb, B, W, H, k, row_offset,
                      (d_aux ? PSH_MODE_FFT : PSH_MODE_EXACT) | flags, d_out_dist,
                      d_out_idx, d_ws, ws_bytes, d_aux, aux_bytes, stream_, &ep);
}

int psh_fft_prepare_embed(const float *d_dataset, int64_t R, int64_t T, int64_t row_stride, int W, int H,
                          const void *d_runs, int nruns, void *d_aux, size_t aux_bytes, void *stream_) {
    if (!d_runs || nruns <= 0) return PSH_E_ARG;
    // spectra, pair norms and twiddles as for the Identity flavour; the energy table holds the embedded
    // energies E2 = sum_n e_n(t)^2 instead of the window energies
    return fft_prepare_impl(d_dataset, R, T, row_stride, W, H, d_aux, aux_bytes, static_cast<const EmbRun *>(d_runs), nruns,
                            (cudaStream_t)stream_);
}

int psh_scan_overflowed(const void *d_ws, int B, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!d_ws || B <= 0) return PSH_E_ARG;
    QState *st = const_cast<QState *>(static_cast<const QState *>(d_ws));
    QState *hst = host_stage(B);
    if (hst == nullptr) return (int)cudaErrorMemoryAllocation;
    PSH_CUDA(cudaMemcpyAsync(hst, st, sizeof(QState) * B, cudaMemcpyDeviceToHost, stream));
    PSH_CUDA(cudaStreamSynchronize(stream));
    int ovf = 0;
    for (int i = 0; i < B; ++i)
        ovf |= (hst[i].overflow != 0 || (hst[i].magic == QSTATE_MAGIC && hst[i].sticky != 0)) ? 1 : 0;
    if (ovf) {  // reported once: the next check starts clean
        clear_sticky_kernel<<<(B + 127) / 128, 128, 0, stream>>>(st, B);
        PSH_LAUNCHED();
    }
    return ovf ? PSH_E_OVERFLOW : PSH_OK;
}

static int merge_impl(const float *d_parts, const int *i_parts, int dstride, int istride, int G, int B, int64_t k,
                      int64_t Tp, float *d_out_dist, int32_t *d_out_idx, int *d_flag, cudaStream_t stream) {
    if (!d_parts || !i_parts || !d_out_dist || !d_out_idx || G <= 0 || B <= 0 || k <= 0 || Tp <= 0) return PSH_E_ARG;
    unsigned long long n = (unsigned long long)G * (unsigned long long)k;
    if (n > 0x7fffffffull) return PSH_E_TOO_LARGE;
    unsigned int npow2 = 1; while (npow2 < n) npow2 <<= 1;
    int use_smem = npow2 <= SORT_SMEM_MAX ? 1 : 0;
    size_t smem = use_smem ? (size_t)npow2 * sizeof(unsigned long long) : 0;
    if (n * 12ull <= MERGE_RANK_SMEM_MAX && !merge_sort_forced()) { use_smem = 2; smem = (size_t)n * 12; }  // merge by rank
    unsigned long long *scratch = nullptr;
    if (!use_smem) PSH_CUDA(cudaMallocAsync(&scratch, (size_t)B * npow2 * sizeof(unsigned long long), stream));
    ProfScope ps_merge(stream, 2);
    merge_kernel<<<B, SEL_THREADS, smem, stream>>>(d_parts, i_parts, dstride, istride, G, B, (unsigned int)k,
                                                   (unsigned long long)Tp, npow2, scratch, use_smem, d_out_dist,
                                                   d_out_idx, d_flag);
    PSH_LAUNCHED();
    if (scratch) PSH_CUDA(cudaFreeAsync(scratch, stream));
    return PSH_OK;
}

int psh_merge_topk(const float *d_dist_parts, const int32_t *d_idx_parts, int G, int B,
                   int64_t k, int64_t Tp, float *d_out_dist, int32_t *d_out_idx, void *stream_) {
    return merge_impl(d_dist_parts, d_idx_parts, 1, 2, G, B, k, Tp, d_out_dist, d_out_idx, nullptr,
                      (cudaStream_t)stream_);
}

int psh_merge_topk_packed(const int32_t *d_rec_parts, int G, int B, int64_t k, int64_t Tp,
                          float *d_out_dist, int32_t *d_out_idx, int32_t *d_overflow_flag, void *stream_) {
    if (!d_rec_parts) return PSH_E_ARG;
    return merge_impl(reinterpret_cast<const float *>(d_rec_parts), d_rec_parts + 1, 3, 3, G, B, k, Tp, d_out_dist,
                      d_out_idx, d_overflow_flag, (cudaStream_t)stream_);
}

// ---- exchange buffers for the peer-memory all-gather (multi-GPU) ----
static size_t xchg_rec_bytes(int G, int B, int64_t k) {   // sized for the LL form: 8 bytes per 4-byte datum
    return align_up((size_t)G * B * (size_t)k * 3 * sizeof(unsigned long long), 256);
}
static bool xchg_ll_enabled() {   // PSH_XCHG_LL=0: always the fence + flag form (A/B measurements)
    const char *e = getenv("PSH_XCHG_LL");
    return !(e != nullptr && e[0] == '0');
}
static unsigned long long xchg_timeout_ns() {   // PSH_XCHG_TIMEOUT_MS: how long a rank waits for its peers (default 30 s)
    const char *e = getenv("PSH_XCHG_TIMEOUT_MS");
    long long ms = e != nullptr ? atoll(e) : 0;
    if (ms <= 0) ms = 30000;
    return (unsigned long long)ms * 1000000ull;
}
static size_t xchg_flag_bytes(int G, int B) { return align_up((size_t)G * B * sizeof(unsigned int), 256); }

size_t psh_xchg_bytes(int G, int B, int64_t k) {
    if (G <= 0 || G > XCHG_MAX_PEERS || B <= 0 || k <= 0) return 0;
    return XCHG_PARITIES * (xchg_rec_bytes(G, B, k) + xchg_flag_bytes(G, B));
}

int psh_xchg_create(size_t bytes, void **d_buf, unsigned char *handle64) {
    if (!d_buf || !handle64 || bytes == 0) return PSH_E_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void *p = nullptr;
    PSH_CUDA(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); cudaGetLastError(); return (int)e; }
    memcpy(handle64, &h, 64);
    *d_buf = p;
    return PSH_OK;
}

int psh_xchg_open(const unsigned char *handle64, void **d_peer) {
    if (!handle64 || !d_peer) return PSH_E_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    cudaError_t e = cudaIpcOpenMemHandle(d_peer, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    return PSH_OK;
}

int psh_xchg_close(void *d_peer) {
    if (!d_peer) return PSH_E_ARG;
    PSH_CUDA(cudaIpcCloseMemHandle(d_peer));
    return PSH_OK;
}

int psh_xchg_destroy(void *d_buf) {
    if (!d_buf) return PSH_E_ARG;
    PSH_CUDA(cudaFree(d_buf));
    return PSH_OK;
}

static int xchg_launch(const int32_t *d_rec_local, void *const *bufs, int G, int rank, int B, int64_t k,
                       int64_t Tp, uint32_t epoch, float *d_out_dist, int32_t *d_out_idx,
                       int32_t *d_flag, int phases, cudaStream_t stream) {
    if (!bufs || G <= 0 || G > XCHG_MAX_PEERS || rank < 0 || rank >= G || B <= 0 || k <= 0 || epoch == 0) return PSH_E_ARG;
    if ((phases & 1) && !d_rec_local) return PSH_E_ARG;
    if ((phases & 2) && (!d_out_dist || !d_out_idx || Tp <= 0)) return PSH_E_ARG;
    unsigned long long n = (unsigned long long)G * (unsigned long long)k;
    if (n > 0x7fffffffull) return PSH_E_TOO_LARGE;
    XchgParams x;
    const size_t rb = xchg_rec_bytes(G, B, k), fb = xchg_flag_bytes(G, B);
    const size_t par = (size_t)(epoch % (unsigned int)XCHG_PARITIES) * (rb + fb);
    for (int g = 0; g < G; ++g) {
        if (!bufs[g]) return PSH_E_ARG;
        unsigned char *base = static_cast<unsigned char *>(bufs[g]) + par;
        x.rec[g] = reinterpret_cast<int *>(base);
        x.flags[g] = reinterpret_cast<unsigned int *>(base + rb);
    }
    x.local_rec = d_rec_local; x.G = G; x.rank = rank; x.epoch = epoch;
    x.timeout_ns = xchg_timeout_ns();
    unsigned int npow2 = 1; while (npow2 < n) npow2 <<= 1;
    int use_smem = npow2 <= SORT_SMEM_MAX ? 1 : 0;
    size_t smem = use_smem ? (size_t)npow2 * sizeof(unsigned long long) : 0;
    if (n * 12ull <= MERGE_RANK_SMEM_MAX && !merge_sort_forced()) { use_smem = 2; smem = (size_t)n * 12; }  // merge by rank
    if (!(phases & 2)) { use_smem = 1; smem = 0; }   // send only: nothing is merged
    if (n * 12ull <= MERGE_RANK_SMEM_MAX && !merge_sort_forced() && xchg_ll_enabled()) {
        // LL form: self-validating 8-byte words, no fence, no flags
        const size_t smem_ll = (phases & 2) ? (size_t)n * 12 : 0;
        {
            ProfScope ps_merge(stream, 2);
            xchg_ll_kernel<<<dim3(G, B), LL_THREADS, smem_ll, stream>>>(x, B, (unsigned int)k, (unsigned long long)Tp,
                                                                         d_out_dist, d_out_idx, d_flag, phases);
        }
        PSH_LAUNCHED();
        return PSH_OK;
    }
    unsigned long long *scratch = nullptr;
    if (!use_smem) PSH_CUDA(cudaMallocAsync(&scratch, (size_t)B * npow2 * sizeof(unsigned long long), stream));
    {
        ProfScope ps_merge(stream, 2);
        xchg_merge_kernel<<<B, SEL_THREADS, smem, stream>>>(x, B, (unsigned int)k, (unsigned long long)Tp, npow2, scratch,
                                                            use_smem, d_out_dist, d_out_idx, d_flag, phases);
    }
    PSH_LAUNCHED();
    if (scratch) PSH_CUDA(cudaFreeAsync(scratch, stream));
    return PSH_OK;
}

int psh_allgather_merge_packed(const int32_t *d_rec_local, void *const *bufs, int G, int rank, int B, int64_t k,
                               int64_t Tp, uint32_t epoch, float *d_out_dist, int32_t *d_out_idx,
                               int32_t *d_flag, void *stream_) {
    return xchg_launch(d_rec_local, bufs, G, rank, B, k, Tp, epoch, d_out_dist, d_out_idx, d_flag, 3,
                       (cudaStream_t)stream_);
}

int psh_xchg_send(const int32_t *d_rec_local, void *const *bufs, int G, int rank, int B, int64_t k,
                  uint32_t epoch, void *stream_) {
    return xchg_launch(d_rec_local, bufs, G, rank, B, k, 1, epoch, nullptr, nullptr, nullptr, 1, (cudaStream_t)stream_);
}

int psh_xchg_merge(void *const *bufs, int G, int rank, int B, int64_t k, int64_t Tp, uint32_t epoch,
                   float *d_out_dist, int32_t *d_out_idx, int32_t *d_flag, void *stream_) {
    return xchg_launch(nullptr, bufs, G, rank, B, k, Tp, epoch, d_out_dist, d_out_idx, d_flag, 2,
                       (cudaStream_t)stream_);
}

int psh_gather_paths(const float *d_dataset, int64_t R, int64_t T, int64_t row_stride,
                     const int32_t *d_idx, int64_t n, int32_t row_offset, int L,
                     float *d_out, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if ((!d_dataset && R > 0) || !d_idx || !d_out || R < 0 || T <= 0 || n < 0 || L <= 0 || L > T || row_stride < T)
        return PSH_E_ARG;  // R == 0: an empty shard, every path is written as zeros
    if (n == 0) return PSH_OK;
    const int warps = 8;
    gather_kernel<<<(unsigned int)((n + warps - 1) / warps), warps * 32, 0, stream>>>(
        d_dataset, R, row_stride, d_idx, n, row_offset, L, d_out);
    PSH_LAUNCHED();
    return PSH_OK;
}

int psh_rv_aggregate(const float *d_paths, const float *d_dist, int B, int64_t k, int L, int H,
                     const int32_t *d_Ts, int nT, float eta, int proba, int vol,
                     float *d_mean, float *d_std, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!d_paths || !d_dist || !d_Ts || !d_mean || !d_std || B <= 0 || k <= 0 || L <= 0 || H <= 0 || H > L)
        return PSH_E_ARG;
    if (nT <= 0 || nT > AGG_MAX_T) return PSH_E_UNSUPPORTED;
    if (proba != 0 && proba != 1) return PSH_E_ARG;
    if (proba == 1 && !(eta > 0.0f)) return PSH_E_ARG;
    rv_aggregate_kernel<<<B, AGG_THREADS, 0, stream>>>(d_paths, d_dist, k, L, H, d_Ts, nT, eta, proba, vol, d_mean,
                                                       d_std);
    PSH_LAUNCHED();
    return PSH_OK;
}

}  // extern "C"
