// pshadow_embed_fft.cuh -- FFT flavour of the EMBEDDED scan (Foveal, PathEmbedding(kernel)),
// included by pshadow.cu behind the FFT section.
//
// The squared embedded distance is a quadratic form of the raw window y_t:
//     S_t = ||ex - K y_t||^2 = ||ex||^2 - 2 g . y_t + ||K y_t||^2,        g = K^T ex  (W values),
// so the cross term is again a correlation of the trajectory with ONE W-vector -- the same
// spectra Z, the same inverse FFT per row pair, the same fft_scan_kernel -- and the quadratic term
// E2[r][t] = sum_n e_n(t)^2 is query-independent: psh_fft_prepare_embed stores it where the
// Identity flavour stores the window energies Y2.  The lower bound
//     LB = ||ex||^2 + E2_t (1 - 16u) - 2 D^_t - slack,
//     slack = 2 cf_u max|FFT(g)| ||y_pair|| + 16u ||ex||^2 + 2u ||g|| ||y_pair||
// (|2 D_t| = 2 |ex . e(t)| <= ||ex||^2 + E2_t, so every rounding of the combination is a few
// u (||ex||^2 + E2_t): the E2 part is taken out of the STORED energies, no per-window work; the
// last term is the rounding of g to fp32) is tested against a threshold widened by 2^-9: the exact embedded evaluation itself (emb_scan_kernel /
// emb_rerank_kernel: box sums good to ~2 ulp, so |s_computed - S| <= gamma S + 2 sqrt(S) eta + eta^2
// with eta <= 8u ||K||_2 ||y_pair||, and 2 sqrt(S) eta <= 2^-10 S + 2^10 eta^2) may differ from the
// true S by that much in either direction.  Survivors are re-evaluated by emb_rerank_kernel with
// the arithmetic of emb_scan_kernel (double-float prefix of the window, runs, s accumulated n
// ascending), so the fft flavour returns what the exact embedded flavour returns up to the last
// bit of a box sum.
#pragma once

// The E2 table is written by fft_prep_energy_kernel<true> (pshadow_fftscan.cuh): fp64 prefix sums, fp64
// box sums, (1 - 16u), scaled by the pair's power of two and rounded DOWN to fp16.

// exact embedded re-rank of the fft filter's candidates: one WARP per candidate window.
// The warp builds the double-float prefix of the window's W samples (lane chunks + fp64 scan of
// the chunk totals, as emb_scan_kernel does for a segment), lanes evaluate the kernel rows
// (row n on lane n % 32, its runs in order: e = fma(c, box, e)), and lane 0 accumulates
// s = fl(s + fl(fl(ex_n - e_n)^2)), n ascending.  grid = (blocks, nq).
constexpr int ERR_WARPS = 8;

__global__ void __launch_bounds__(ERR_WARPS * 32) emb_rerank_kernel(const float *__restrict__ ds, long long row_stride,
                                                                    unsigned int Tp, int W, int d,
                                                                    const float *__restrict__ qemb,
                                                                    const EmbRun *__restrict__ runs, int nruns,
                                                                    QState *st_all, const unsigned int *__restrict__ cand_all,
                                                                    unsigned long long *keys_all, unsigned int cap) {
    extern __shared__ __align__(16) unsigned char er_smem[];
    const int dpad = (d + 3) & ~3;
    float *exs = reinterpret_cast<float *>(er_smem);                       // (dpad)
    EmbRun *runs_s = reinterpret_cast<EmbRun *>(exs + dpad);               // (nruns)
    float2 *ps_all = reinterpret_cast<float2 *>(runs_s + nruns);           // (warps, W + 2)
    float *es_all = reinterpret_cast<float *>(ps_all + (size_t)ERR_WARPS * (W + 2));   // (warps, dpad)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2 *Ps = ps_all + (size_t)warp * (W + 2);
    float *es = es_all + (size_t)warp * dpad;

    const int b = blockIdx.y;
    QState *st = st_all + b;
    const unsigned int craw = st->ccount;
    const unsigned int C = min(craw, cap);
    if (craw > cap && threadIdx.x == 0 && blockIdx.x == 0) { st->overflow = 1; st->sticky = 1; }
    if (blockIdx.x * ERR_WARPS >= C) return;
    for (int j = threadIdx.x; j < d; j += ERR_WARPS * 32) exs[j] = qemb[(size_t)b * d + j];
    for (int j = threadIdx.x; j < nruns; j += ERR_WARPS * 32) runs_s[j] = runs[j];
    __syncthreads();
    const float s_thr = st->s_thr, qn = st->qnorm;
    const unsigned long long tau = st->tau_key;
    const unsigned int *cand = cand_all + (size_t)b * cap;
    unsigned long long *dst = keys_all + ((size_t)b * 2 + st->cur) * cap;
    const int epl = (W + 31) / 32;

    for (unsigned int c = blockIdx.x * ERR_WARPS + warp; c < C; c += gridDim.x * ERR_WARPS) {
        const unsigned int flat = cand[c];
        const unsigned int r = flat / Tp, t = flat - r * Tp;
        const float *y = ds + (long long)r * row_stride + t;
        // double-float prefix of the window: Ps[i] = sum_{j<i} y_j
        const int e0 = lane * epl;
        float hi = 0.0f, lo = 0.0f;
        for (int i = 0; i < epl; ++i) df_add(hi, lo, e0 + i < W ? __ldg(y + e0 + i) : 0.0f);
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
            if (e < W) {
                df_add(hi, lo, __ldg(y + e));
                const float h2 = __fadd_rn(hi, lo);
                Ps[e + 1] = make_float2(h2, __fsub_rn(lo, __fsub_rn(h2, hi)));
            }
        }
        __syncwarp();
        // rows: lane l owns rows l, l+32, ...; the runs of a row are contiguous and ordered
        float e = 0.0f;
        for (int rr = 0; rr < nruns; ++rr) {
            const EmbRun rn = runs_s[rr];
            if ((rn.row & 31) == lane) {
                const float2 pa = Ps[rn.a], pb = Ps[rn.b];
                const float box = __fadd_rn(__fsub_rn(pb.x, pa.x), __fsub_rn(pb.y, pa.y));
                e = fmaf(rn.c, box, e);
                if (rr + 1 == nruns || runs_s[rr + 1].row != rn.row) { es[rn.row] = e; e = 0.0f; }
            }
        }
        __syncwarp();
        if (lane == 0) {
            float s = 0.0f;
            int prev = -1;
            for (int rr = 0; rr < nruns; ++rr) {   // rows that have runs, ascending (rows without runs add 0)
                const int n = runs_s[rr].row;
                if (n == prev) continue;
                prev = n;
                const float df = __fsub_rn(exs[n], es[n]);
                s = __fadd_rn(s, __fmul_rn(df, df));
            }
            if (s <= s_thr) {
                const unsigned long long key = ((unsigned long long)__float_as_uint(dist_from_s(s, qn)) << 32) | flat;
                if (key <= tau) {
                    const unsigned int pos = atomicAdd(&st->count, 1u);
                    if (pos < cap) dst[pos] = key;
                }
            }
        }
        __syncwarp();   // Ps / es are reused by the warp's next candidate
    }
}
