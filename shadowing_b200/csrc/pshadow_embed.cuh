// pshadow_embed.cuh -- scan in EMBEDDED space (any linear PathEmbedding: Foveal, dense kernels),
// included by pshadow.cu.
//
// Reference: PathEmbedding.forward = conv1d with a (d, 1, W) kernel zero-padded by the horizon
// (path_embedding.py:117-132, 48-51; Foveal :142-172) materialises e_n(t) = sum_j K[n][j] y[t+j] for
// every window, then RelativeMSE over the d dimensions (path_distance.py:62-65).  Here nothing is
// materialised.  The host decomposes every kernel row into RUNS of equal taps,
//     e_n(t) = sum_runs c * (P[t + b] - P[t + a]),        P[i] = sum_{j<i} y[t0 + j],
// (Foveal: ONE run per row, b = W: 34 box sums per window instead of 34 x 126 multiply-adds), and
// each warp builds the prefix P of its staged segment ONCE, in double-float arithmetic (hi + lo
// pairs of fp32: ~48 bits, so a box sum is good to ~2 ulp whatever the cancellation), then every
// lane evaluates its 12 windows (t = lane + 32 i: consecutive lanes read consecutive float2's, no
// bank conflicts) row by row for up to QG queries at a time.  The squared distance accumulates
// like the reference's: s = fl(s + fl(fl(ex_n - e_n)^2)), n ascending; thresholds, candidate
// lists, select and finalise are the exact flavour's.
#pragma once

struct EmbRun { int row; int a; int b; float c; };   // 0 <= a < b <= W

struct EmbParams {
    const EmbRun *runs;
    int nruns;
    int d;          // embedding dimension (queries are (nq, d), staged with stride sp.wpad)
    int ps_n;       // float2 entries of a warp's prefix array (>= SEG + W, even)
    // fft flavour of the embedded scan (NULL / 0: exact flavour only), see pshadow_embed_fft.cuh
    const float *g;   // (nq, W) cross-term vectors g = K^T ex
};

constexpr int EMB_QG = 3;   // queries evaluated per pass over a staged segment

__device__ __forceinline__ void df_add(float &hi, float &lo, float y) {
    // (hi, lo) += y, error-free transformation (Knuth TwoSum); lo collects the rounding errors
    const float s = __fadd_rn(hi, y);
    const float bb = __fsub_rn(s, hi);
    const float err = __fadd_rn(__fsub_rn(hi, __fsub_rn(s, bb)), __fsub_rn(y, bb));
    hi = s;
    lo = __fadd_rn(lo, err);
}

template <int QG>
__global__ void __launch_bounds__(SCAN_THREADS, 2) emb_scan_kernel(const ScanParams p, const EmbParams ep) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *exs = reinterpret_cast<float *>(smem_raw);                               // (nq, wpad)
    EmbRun *runs = reinterpret_cast<EmbRun *>(exs + (size_t)p.nq * p.wpad);          // (nruns)
    float2 *ps_all = reinterpret_cast<float2 *>(runs + ep.nruns);                    // (warps, ps_n)
    float *bufs = reinterpret_cast<float *>(ps_all + (size_t)SCAN_WARPS * ep.ps_n);  // (warps, 2, buf_floats)
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(bufs + (size_t)SCAN_WARPS * 2 * p.buf_floats);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *mybuf = bufs + (size_t)warp * 2 * p.buf_floats;
    float2 *Ps = ps_all + (size_t)warp * ep.ps_n;
    const uint32_t bar0 = smem_u32(&bars[warp * 2]);

    for (int i = threadIdx.x; i < p.nq * p.wpad; i += SCAN_THREADS) {
        const int b = i / p.wpad, j = i - b * p.wpad;
        exs[i] = j < ep.d ? p.queries[(size_t)b * ep.d + j] : 0.0f;
    }
    for (int i = threadIdx.x; i < ep.nruns; i += SCAN_THREADS) runs[i] = ep.runs[i];
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

    const int need = SEG + p.W - 1;             // samples a full segment touches
    const int epl = ((need + 31) / 32) | 1;     // samples per lane in the prefix pass (odd: no bank conflicts)
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
            __syncwarp();
        }

        // ---- prefix sums of the staged samples in double-float arithmetic ----
        // pass 1: this lane's chunk total; warp scan of the totals in fp64; pass 2: continue the
        // accumulation from the lane's exclusive offset and store every partial sum (hi, lo)
        const int e0 = lane * epl;
        {
            float hi = 0.0f, lo = 0.0f;
            for (int i = 0; i < epl; ++i) {
                const int e = e0 + i;
                df_add(hi, lo, e < tk.nvalid ? buf[e] : 0.0f);
            }
            const double tot = (double)hi + (double)lo;
            double incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double u = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += u;
            }
            const double off = incl - tot;
            hi = (float)off;
            lo = (float)(off - (double)hi);
            if (lane == 0) Ps[0] = make_float2(0.0f, 0.0f);
            for (int i = 0; i < epl; ++i) {
                const int e = e0 + i;
                if (e < need) {
                    df_add(hi, lo, e < tk.nvalid ? buf[e] : 0.0f);
                    const float h2 = __fadd_rn(hi, lo);
                    Ps[e + 1] = make_float2(h2, __fsub_rn(lo, __fsub_rn(h2, hi)));
                }
            }
        }
        __syncwarp();

        const int t0 = tk.t0;
        const unsigned int flat0 =
            (unsigned int)((unsigned long long)tk.row * (unsigned long long)p.Tp + (unsigned long long)(t0 + lane));
        const float2 *Pl = Ps + lane;   // window i of this lane: local index lane + 32 i
        float2 pend[WPT];
#pragma unroll
        for (int i = 0; i < WPT; ++i) pend[i] = Pl[p.W + 32 * i];

        for (int g0 = 0; g0 < p.nq; g0 += QG) {
            float acc[QG][WPT], e[WPT];
#pragma unroll
            for (int q = 0; q < QG; ++q)
#pragma unroll
                for (int i = 0; i < WPT; ++i) acc[q][i] = 0.0f;
#pragma unroll
            for (int i = 0; i < WPT; ++i) e[i] = 0.0f;
#pragma unroll 1
            for (int r = 0; r < ep.nruns; ++r) {
                const EmbRun run = runs[r];
                const float2 *pa = Pl + run.a;
                if (run.b == p.W) {   // warp-uniform: trailing box (every Foveal row)
#pragma unroll
                    for (int i = 0; i < WPT; ++i) {
                        const float2 a = pa[32 * i];
                        const float box = __fadd_rn(__fsub_rn(pend[i].x, a.x), __fsub_rn(pend[i].y, a.y));
                        e[i] = fmaf(run.c, box, e[i]);
                    }
                } else {
                    const float2 *pb = Pl + run.b;
#pragma unroll
                    for (int i = 0; i < WPT; ++i) {
                        const float2 a = pa[32 * i], bq = pb[32 * i];
                        const float box = __fadd_rn(__fsub_rn(bq.x, a.x), __fsub_rn(bq.y, a.y));
                        e[i] = fmaf(run.c, box, e[i]);
                    }
                }
                const bool row_done = (r + 1 == ep.nruns) || (runs[r + 1].row != run.row);
                if (row_done) {   // warp-uniform
#pragma unroll
                    for (int q = 0; q < QG; ++q) {
                        const float exq = exs[(size_t)min(g0 + q, p.nq - 1) * p.wpad + run.row];
#pragma unroll
                        for (int i = 0; i < WPT; ++i) {
                            const float df = __fsub_rn(exq, e[i]);
                            acc[q][i] = __fadd_rn(acc[q][i], __fmul_rn(df, df));
                        }
                    }
#pragma unroll
                    for (int i = 0; i < WPT; ++i) e[i] = 0.0f;
                }
            }

            // ---- epilogue: windows that beat the running threshold join the query's key list ----
#pragma unroll
            for (int q = 0; q < QG; ++q) {
                const int b = g0 + q;
                if (b >= p.nq) break;   // warp-uniform
                const float s_thr = ld_volatile_f32(&p.st[b].s_thr);
                unsigned int mask = 0;
#pragma unroll
                for (int i = 0; i < WPT; ++i)
                    if (acc[q][i] <= s_thr && t0 + lane + 32 * i < tk.tp_eff) mask |= 1u << i;
                if (!__any_sync(FULL, mask != 0)) continue;
                unsigned long long key[WPT];
                const float qn = p.st[b].qnorm;
                const unsigned long long tau = ld_volatile_u64(&p.st[b].tau_key);
#pragma unroll
                for (int i = 0; i < WPT; ++i) {
                    key[i] = 0;
                    if (mask & (1u << i)) {
                        const float dd = dist_from_s(acc[q][i], qn);
                        key[i] = ((unsigned long long)__float_as_uint(dd) << 32) | (unsigned long long)(flat0 + 32u * i);
                        if (key[i] > tau) mask &= ~(1u << i);
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
                    if (lane == 31) base = atomicAdd(&p.st[b].count, (unsigned int)total);
                    base = __shfl_sync(FULL, base, 31);
                    unsigned int pos = base + (unsigned int)(incl - cnt);
                    unsigned long long *dst = p.keys + ((size_t)b * 2 + p.st[b].cur) * p.cap;
#pragma unroll
                    for (int i = 0; i < WPT; ++i)
                        if (mask & (1u << i)) {
                            if (pos < p.cap) dst[pos] = key[i];
                            ++pos;
                        }
                }
            }
        }
        __syncwarp();  // every lane is done with buf and Ps before they are refilled
        if (!have_next) break;
        tk = tn;
        task = next;
    }
}
