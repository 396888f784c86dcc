// pshadow_fftscan.cuh -- FFT flavour: dataset preparation, query preparation, the scan kernel.
// Included by pshadow.cu behind pshadow_fft.cuh / pshadow_fft2.cuh / pshadow_embed.cuh (see pshadow_fft2.cuh
// for the data formats and the arithmetic).
#pragma once

struct FftAux {            // device pointers into the caller's aux buffer
    float2 *tw32;          // exp(+2 pi i m / 4096), m < 4096
    double2 *tw64;
    float4 *pinfo;         // (npairs) { ||y_pair||_2 rounded up, zqerr, es, m2 = -2 es / zs }
    __half2 *Z;            // (npairs, 4096) spectra of y_a + i y_b, scaled by zs (power of two), fp16 pairs
    __half2 *Y2;           // (npairs, 4096) (Y2_a, Y2_b)[pos] * es rounded down to fp16, scan output order, +inf = no window
    long long npairs;
    int nsegv, hop, span;  // virtual rows, see ScanParams
    long long VR;
    size_t total;
    int nfft;              // transform length: 1024 (one warp per transform, pshadow_fft3.cuh) or 4096 (one CTA)
    int ncy;               // nfft = 1024: 128-window groups of a piece's energy row that hold windows
    float4 *tw2;           // nfft = 1024: {w^(lane 2j), w^(lane (2j+1))} at [j * 32 + lane], w = exp(2 pi i / 1024)
};

// Transform length of the fft flavour for a context of W samples: 1024-point pieces while a piece still
// yields enough windows (1025 - W of 1024), the 4096-point transform beyond.  PSH_FFT_N=4096 / 1024 forces
// one (A/B measurements, tests); prepare and scan must see the same setting.
constexpr int FFT3_MAX_W = 384;
inline int fft_length_for(int W) {
    const char *e = getenv("PSH_FFT_N");
    const int forced = (e != nullptr && e[0] != 0) ? atoi(e) : 0;
    if (forced == 4096) return 4096;
    if (forced == 1024 && W <= 768) return 1024;
    return W <= FFT3_MAX_W ? 1024 : 4096;
}

inline bool fft_aux_layout(long long R, long long T, int W, int H, unsigned char *base, FftAux &a, int nfft = 0) {
    if (R <= 0 || T <= 0 || W <= 0 || W > fftx::N / 2 || H < 0 || T - W - H + 1 <= 0) return false;
    const long long Tp = T - W - H + 1;
    a.nfft = (nfft == 1024 || nfft == 4096) ? nfft : fft_length_for(W);   // (a scan takes the length its aux was prepared with)
    if (T <= a.nfft) { a.nsegv = 1; a.hop = a.nfft; a.span = (int)Tp; }
    else {
        a.hop = (a.nfft - W + 1) & ~3;                    // windows per piece
        a.span = a.hop;
        a.nsegv = (int)((Tp + a.hop - 1) / a.hop);
    }
    a.ncy = (a.span + 127) / 128;
    a.VR = R * (long long)a.nsegv;
    if (a.VR > 0x7fffffffLL) return false;
    a.npairs = (a.VR + 1) / 2;
    size_t off = 0;
    a.tw32 = reinterpret_cast<float2 *>(base + off); off += sizeof(float2) * fftx::N;
    a.tw64 = reinterpret_cast<double2 *>(base + off); off += sizeof(double2) * fftx::N;
    a.tw2 = reinterpret_cast<float4 *>(base + off); off += sizeof(float4) * 512;
    a.pinfo = reinterpret_cast<float4 *>(base + off); off += (sizeof(float4) * (size_t)a.npairs + 255) / 256 * 256;
    a.Z = reinterpret_cast<__half2 *>(base + off); off += sizeof(__half2) * (size_t)a.nfft * (size_t)a.npairs;
    a.Y2 = reinterpret_cast<__half2 *>(base + off); off += sizeof(__half2) * (size_t)a.nfft * (size_t)a.npairs;
    a.total = off;
    return true;
}

// debug / test entry: batched 4096-point transform of n independent signals
//   dir -1 / +1: table-twiddle forward / inverse (the dataset spectra use the forward one)
//   dir +3     : the scan's packed inverse, output un-permuted
__global__ void __launch_bounds__(fftx::THREADS) fft_debug_kernel(const float2 *__restrict__ in, float2 *out,
                                                                  const float2 *__restrict__ tw, int dir) {
    __shared__ float2 ex[fx2::EX_FLOAT2];
    const int tid = threadIdx.x;
    const float2 *x = in + (size_t)blockIdx.x * fftx::N;
    float2 *y = out + (size_t)blockIdx.x * fftx::N;
    float2 v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = x[tid + 256 * i];
    if (dir >= 3) {
        fx2::ifft4096(v, ex, tid, fx2::load_seeds(tw, tid), []() {}, []() {});
        const int kb = fx2::out_base(tid);
#pragma unroll
        for (int c = 0; c < 16; ++c) y[kb + 256 * c] = v[c];
        return;
    }
    const fftx::TwSeeds seeds = fftx::load_seeds(tw, tid);
    if (dir > 0) fftx::fft4096<1, true>(v, ex, tw, tid, seeds);
    else fftx::fft4096<-1, true>(v, ex, tw, tid, seeds);
#pragma unroll
    for (int c = 0; c < 16; ++c) y[tid + 256 * c] = v[c];
}

__device__ __forceinline__ float block_max_256(float v, float *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float m = red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    return m;
}
__device__ __forceinline__ double block_sum_256(double v, double *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i];
    return s;
}
// power of two s with m * s in [2^(top-1), 2^top); 1 for m = 0 / non-finite; exponent clamped to +-100
__device__ __forceinline__ float pow2_scale(float m, int top) {
    if (!(m > 0.0f) || !(m < __int_as_float(0x7f800000))) return 1.0f;
    int e;
    frexpf(m, &e);                 // m = f 2^e, f in [0.5, 1)
    int sh = top - e;
    sh = sh < -100 ? -100 : (sh > 100 ? 100 : sh);
    return ldexpf(1.0f, sh);
}

// spectra of row pairs: fp32 transform (table twiddles), quantised to fp16 pairs with the pair's scale;
// pair norm; MEASURED quantisation error.  One CTA per pair.
__global__ void __launch_bounds__(fftx::THREADS) fft_prep_spectra_kernel(const float *__restrict__ ds, int T,
                                                                         long long row_stride, FftAux a) {
    __shared__ float2 ex[fftx::EX_FLOAT2];
    __shared__ double red[8];
    __shared__ float redf[8];
    const int tid = threadIdx.x;
    const long long pair = blockIdx.x;
    const long long va = 2 * pair, vb = va + 1;
    const bool has_b = vb < a.VR;
    const long long rowa = va / a.nsegv, rowb = (has_b ? vb : va) / a.nsegv;
    const int oa = (int)(va - rowa * a.nsegv) * a.hop, ob = (int)((has_b ? vb : va) - rowb * a.nsegv) * a.hop;
    const float *ya = ds + rowa * row_stride + oa;
    const float *yb = ds + rowb * row_stride + ob;
    float2 v[16];
    double e = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int n = tid + 256 * i;
        const float xa = oa + n < T ? ya[n] : 0.0f;
        const float xb = (ob + n < T && has_b) ? yb[n] : 0.0f;
        v[i] = make_float2(xa, xb);
        e += (double)xa * (double)xa + (double)xb * (double)xb;
    }
    fftx::fft4096<-1, true>(v, ex, a.tw32, tid, fftx::load_seeds(a.tw32, tid));
    const double etot = block_sum_256(e, red);
    float mx = 0.0f;
#pragma unroll
    for (int c = 0; c < 16; ++c) mx = fmaxf(mx, fmaxf(fabsf(v[c].x), fabsf(v[c].y)));
    const float zs = pow2_scale(block_max_256(mx, redf), 14);   // largest component in [2^13, 2^14)
    const double inv = 1.0 / (double)zs;
    __half2 *z = a.Z + (size_t)pair * fftx::N;
    double err = 0.0;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        const float sx = fminf(fmaxf(v[c].x * zs, -65504.0f), 65504.0f);
        const float sy = fminf(fmaxf(v[c].y * zs, -65504.0f), 65504.0f);
        const __half2 h = __floats2half2_rn(sx, sy);
        const float2 f = __half22float2(h);
        const double dx = (double)f.x * inv - (double)v[c].x, dy = (double)f.y * inv - (double)v[c].y;
        err += dx * dx + dy * dy;
        z[tid + 256 * c] = h;
    }
    err = block_sum_256(err, red);
    if (tid == 0) {
        float4 pi;
        pi.x = __double2float_ru(sqrt(etot) * (1.0 + 1e-7));                 // ||y_pair||
        pi.y = __double2float_ru(sqrt(err / (double)fftx::N) * (1.0 + 1e-6)); // ||Z^ - Z||_2 / sqrt(N)
        pi.z = 0.0f;                                                          // es: fft_prep_energy_kernel
        pi.w = zs;                                                            // (replaced by m2 there)
        a.pinfo[pair] = pi;
    }
}

// window energies of a pair's two virtual rows, in the scan's output order (position p = tid + 256 c holds
// window t = kb(tid) + 256 c, kb = tid's hex digits swapped), scaled by the pair's power of two es and rounded DOWN to fp16; +inf
// where the virtual row owns no window.  Identity: Y2[t] = sum_{j<W} y_{t+j}^2 from an fp64 prefix sum.
// EMBK (pshadow_embed_fft.cuh): E2[t] = sum_n e_n(t)^2 (1 - 16u) from an fp64 prefix sum of y and the
// kernel's runs.  One CTA per pair.
template <int NFFT> __device__ __forceinline__ int fft_window_of_pos(int pos);
template <> __device__ __forceinline__ int fft_window_of_pos<4096>(int pos) { return fx2::out_base(pos & 255) + (pos & ~255); }
template <> __device__ __forceinline__ int fft_window_of_pos<1024>(int pos) { const int j = pos & 3, l = (pos >> 2) & 31, g = pos >> 7; return l + 32 * (4 * g + j); }

template <bool EMBK, int NFFT>
__global__ void __launch_bounds__(fftx::THREADS) fft_prep_energy_kernel(const float *__restrict__ ds, int T,
                                                                        long long row_stride, int W, int Tp, FftAux a,
                                                                        const EmbRun *__restrict__ runs, int nruns) {
    constexpr int PER = NFFT / fftx::THREADS;   // samples (and output positions) per thread
    __shared__ double pfx[NFFT + 1];
    __shared__ double wsum[8];
    __shared__ float redf[8];
    extern __shared__ EmbRun runs_s[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (EMBK)
        for (int i = tid; i < nruns; i += fftx::THREADS) runs_s[i] = runs[i];
    const long long pair = blockIdx.x;
    float en[2][PER];
    float mx = 0.0f;
    const float INF = __int_as_float(0x7f800000);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const long long vr = 2 * pair + h;
        const bool have = vr < a.VR;
        const long long row = (have ? vr : 2 * pair) / a.nsegv;
        const int piece = (int)((have ? vr : 2 * pair) - row * a.nsegv);
        const int o0 = piece * a.hop;                     // first sample / window of this virtual row
        const float *y = ds + row * row_stride + o0;
        double loc[PER];
        double run = 0.0;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int n = PER * tid + i;
            const double x = o0 + n < T ? (double)y[n] : 0.0;
            run += EMBK ? x : x * x;
            loc[i] = run;
        }
        double incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double u = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += u;
        }
        __syncthreads();                                  // pfx / wsum of the other row are no longer read
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        double off = incl - run;
        for (int w = 0; w < warp; ++w) off += wsum[w];
        if (tid == 0) pfx[0] = 0.0;
#pragma unroll
        for (int i = 0; i < PER; ++i) pfx[PER * tid + i + 1] = off + loc[i];
        __syncthreads();
#pragma unroll
        for (int c = 0; c < PER; ++c) {
            const int pos = tid + 256 * c;
            const int t = fft_window_of_pos<NFFT>(pos);   // the window the scan's epilogue finds at position pos
            const bool mine = have && t < a.span && o0 + t < Tp;   // the windows this virtual row owns
            float out = INF;
            if (mine) {
                if (EMBK) {
                    double e2 = 0.0, e = 0.0;
                    for (int r = 0; r < nruns; ++r) {
                        const EmbRun rn = runs_s[r];
                        e += (double)rn.c * (pfx[t + rn.b] - pfx[t + rn.a]);
                        if (r + 1 == nruns || runs_s[r + 1].row != rn.row) { e2 += e * e; e = 0.0; }
                    }
                    out = __double2float_rd(e2 * (1.0 - 16.0 * 5.9604644775390625e-8));   // 16u E2 rounding allowance
                } else {
                    out = __double2float_rd(pfx[t + W] - pfx[t]);
                }
                out = fmaxf(out, 0.0f);
                if (out < INF) mx = fmaxf(mx, out);
            }
            en[h][c] = out;
        }
    }
    const float zs = a.pinfo[pair].w;
    float es = pow2_scale(block_max_256(mx, redf), 15);   // largest energy in [2^14, 2^15)
    float m2 = -2.0f * es / zs;
    if (!(fabsf(m2) < __int_as_float(0x7f800000)) || fabsf(m2) < 1e-30f) { es = zs; m2 = -2.0f; }
    __half2 *o = a.Y2 + (size_t)pair * NFFT;
#pragma unroll
    for (int c = 0; c < PER; ++c) {
        // a finite energy stays finite (a clamped value is still a lower bound); +inf marks "no window"
        const float ea = en[0][c] < INF ? fminf(en[0][c] * es, 65504.0f) : INF;
        const float eb = en[1][c] < INF ? fminf(en[1][c] * es, 65504.0f) : INF;
        if (NFFT == 1024) {
            // 1024-point flavour: bf16 pairs (the scan unpacks them on the integer pipe); the values are >= 0,
            // so dropping the low 16 bits rounds DOWN; +inf stays +inf
            const unsigned int w = (__float_as_uint(ea) >> 16) | (__float_as_uint(eb) & 0xffff0000u);
            reinterpret_cast<unsigned int *>(o)[tid + 256 * c] = w;
        } else {
            o[tid + 256 * c] = __halves2half2(__float2half_rd(ea), __float2half_rd(eb));
        }
    }
    if (tid == 0) {
        a.pinfo[pair].z = es;
        a.pinfo[pair].w = m2;
    }
}

// ------------------------------------------------------------------------------------------
// per-query state of the fft flavour in the workspace: threshold histogram + published threshold
// ------------------------------------------------------------------------------------------
// Logarithmic bins on the float's bit pattern: bits >> 13 (8 exponent + 10 mantissa bits, 0.1 % wide),
// HB fine bins = 8 binades below 8 Q2, and HC coarse bins of 128 fine bins each (a second counter per
// entry).  Every entry is the UPPER bound UB of the exact squared distance of ONE distinct window; k
// entries at or below a bin => the k-th exact squared distance <= the bin's upper edge => filtering with
// edge * widen2 loses nothing.  No range estimate, no scale agreed between CTAs; ONE warp re-derives the
// threshold with two dependent rounds of loads (coarse counts, then the fine bins of one coarse bin).
constexpr int HB = 8192;
constexpr int HC = 64;
constexpr int HFINE_PER_COARSE = HB / HC;   // 128
constexpr int H_THR = HB + HC;              // published threshold (float bits, atomicMin)
// (each in a 128-byte line of its own: the arrivals, the slot draws and the polls of `done` come from every
// CTA / warp of the launch and must not queue up behind one another in one L2 line)
constexpr int H_ARR = HB + HC + 32;         // seed arrivals of the launch (query 0's slot)
constexpr int H_SLOT = HB + HC + 64;        // dynamic pair-slot counter of the launch (query 0's slot)
constexpr int H_DONE = HB + HC + 96;        // the launch's seed thresholds have been published (query 0's slot)
constexpr int HSTRIDE = HB + HC + 128;      // uints per query

__device__ __forceinline__ int hist_base(float q2) { return (int)(__float_as_uint(8.0f * q2) >> 13) - HB; }
// one entry (`n` of them in the same bin) for the upper bound ub >= 0
__device__ __forceinline__ void hist_add_ub(unsigned int *hq, int bin, unsigned int n) {
    atomicAdd(hq + bin, n);
    atomicAdd(hq + HB + bin / HFINE_PER_COARSE, n);
}

// conj(FFT_4096(g padded))/4096 per query (direct fp64 DFT on the exact twiddle table), max_k |G_k|, and
// the query state: ||q|| in torch's contiguous-reduction order, Q2, ||g||; zeroes the query's histogram.
// g is the correlated vector: the context itself (Identity) or K^T ex (embedded scan, q = ex).
// grid = (4096/16, nq): 16 frequencies per CTA, 16 threads share one frequency (256 CTAs: the whole
// GPU works on one query's 10^6 fp64 multiply-adds; with 64 CTAs it took 13 us).
constexpr int QFFT_K = 16;
constexpr int QFFT_PARTS = 16;
constexpr int QFFT_THREADS = QFFT_K * QFFT_PARTS;   // 256
constexpr int QMAXP = fftx::N / QFFT_K;             // per-query partial maxima of |FFT(q)|, one per CTA

__device__ __forceinline__ void qstate_init(const float *__restrict__ x, int n, QState *st, float gnorm) {
    // one warp; lanes 0..7 own torch's 8 interleaved partial sums (path_distance.py:65 `x.norm(dim=-1)`)
    const int lane = threadIdx.x & 31;
    const int n8 = (n / 8) * 8;
    float acc = 0.0f;
    if (lane < 8)
        for (int j = lane; j < n8; j += 8) acc = __fadd_rn(acc, __fmul_rn(x[j], x[j]));
    float s = 0.0f;
#pragma unroll
    for (int l = 0; l < 8; ++l) s = __fadd_rn(s, __shfl_sync(FULL, acc, l));
    for (int j = n8; j < n; ++j) s = __fadd_rn(s, __fmul_rn(x[j], x[j]));  // scalar tail (all lanes alike)
    double q2 = 0.0;
    for (int j = lane; j < n; j += 32) q2 += (double)x[j] * (double)x[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q2 += __shfl_xor_sync(FULL, q2, o);
    if (lane == 0) {
        QState z;
        z.sticky = (st->magic == QSTATE_MAGIC) ? st->sticky : 0u;
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
        z.gnorm = gnorm;
        for (int i = 0; i < 2; ++i) z.pad[i] = 0;
        *st = z;
    }
}

__global__ void __launch_bounds__(QFFT_THREADS) qfft_kernel(const float *__restrict__ q, int qlen,
                                                            const float *__restrict__ g, int W,
                                                            const double2 *__restrict__ tw64, float2 *Qc, QState *st,
                                                            unsigned int *hist, float *qmaxp, int nfft) {
    extern __shared__ double qd[];
    __shared__ double red[QFFT_THREADS / 32];
    const int b = blockIdx.y, tid = threadIdx.x;
    for (int j = tid; j < W; j += QFFT_THREADS) qd[j] = (double)g[(size_t)b * W + j];
    // this query's threshold histogram starts at zero, its published threshold at +inf
    unsigned int *hq = hist + (size_t)b * HSTRIDE;
    const int zb = HB / (int)gridDim.x;   // fine bins this CTA clears (gridDim.x = nfft / 16 divides HB)
    for (int i = tid; i < zb; i += QFFT_THREADS) hq[blockIdx.x * zb + i] = 0u;
    if (blockIdx.x == 0 && tid < HSTRIDE - HB) hq[HB + tid] = (HB + tid == H_THR) ? 0x7f800000u : 0u;
    __syncthreads();
    if (blockIdx.x == 0 && tid < 32) {   // ||g||_2 of the correlated vector, rounded up; the query state
        double s2 = 0.0;
        for (int j = tid; j < W; j += 32) s2 += qd[j] * qd[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(FULL, s2, o);
        qstate_init(q + (size_t)b * qlen, qlen, st + b, __double2float_ru(sqrt(s2) * (1.0 + 1e-7)));
    }
    const int k = blockIdx.x * QFFT_K + tid / QFFT_PARTS, part = tid % QFFT_PARTS;
    const int per = (W + QFFT_PARTS - 1) / QFFT_PARTS;
    const int j0 = part * per, j1 = min(W, j0 + per);
    // G_k = sum_j g_j exp(-2 pi i j k / N): the twiddle of the thread's first term and the step exp(2 pi i k / N)
    // come from the exact table, the others by recurrence (<= 16 complex fp64 products: ~1e-15, far below the
    // fp32 rounding of the result) -- a gather of the table per term made this kernel latency-bound (12 us)
    double re = 0.0, im = 0.0;
    const int tws = fftx::N / nfft;       // the table holds exp(2 pi i m / 4096)
    double2 w = __ldg(tw64 + ((j0 * k * tws) & (fftx::N - 1)));
    const double2 w1 = __ldg(tw64 + ((k * tws) & (fftx::N - 1)));
    for (int j = j0; j < j1; ++j) {
        re += qd[j] * w.x;
        im -= qd[j] * w.y;
        const double wx = w.x * w1.x - w.y * w1.y;
        w.y = w.x * w1.y + w.y * w1.x;
        w.x = wx;
    }
#pragma unroll
    for (int o = 1; o < QFFT_PARTS; o <<= 1) {
        re += __shfl_xor_sync(FULL, re, o);
        im += __shfl_xor_sync(FULL, im, o);
    }
    // (1024-point flavour: the order in which a lane of the scan fetches its 32 values, see pshadow_fft3.cuh)
    if (part == 0) Qc[(size_t)b * fftx::N + (nfft == fftx::N ? k : ((((k >> 5) >> 1) * 32 + (k & 31)) * 2 + ((k >> 5) & 1)))] =
        make_float2((float)(re / nfft), (float)(-im / nfft));
    double mx = re * re + im * im;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(FULL, mx, o));
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    if (tid == 0) {
        for (int i = 1; i < QFFT_THREADS / 32; ++i) mx = fmax(mx, red[i]);
        qmaxp[(size_t)b * QMAXP + blockIdx.x] = __double2float_ru(sqrt(mx) * (1.0 + 1e-7));
    }
}

struct FftScanParams {
    const __half2 *Z;
    const __half2 *Y2;
    const float4 *pinfo;
    const float2 *tw;
    const float2 *Qc;      // (nq, 4096)
    const float *qmaxp;    // (nq, QMAXP)
    int Tp, nq;
    int nsegv, hop;
    int npairs;
    long long VR;
    int i0, i1;            // pair slots of this launch
    long long perm;        // pair = (slot * perm) mod npairs, gcd(perm, npairs) = 1
    double inv_np;
    QState *st;
    unsigned int *cand;
    unsigned int cap;
    float cf_u;            // CF * 2^-24: |c^_t - c_t| <= cf_u * Qmax * ynorm  (CF = 512, theory ~165)
    unsigned int *hist;    // (nq, HSTRIDE)
    unsigned int k;
    float widen2;          // (1 + 2 (W+8) u)^2 (1 + 1e-6): exact-sequence rounding, both directions
    int seed;              // 1: the launch seeds its own threshold from its first pair per CTA (thresholds start at +inf)
    unsigned int seed_need;  // seed arrivals a CTA waits for before it derives its first threshold
    unsigned int refresh_mask;  // a CTA re-derives the thresholds when ((iteration + blockIdx) & mask) == 0
    // Identity: 12u (Q2 + ynorm^2); embedded scan (pshadow_embed_fft.cuh): slack_coef Q2 + g_coef ||g|| ynorm
    float slack_coef, g_coef;
    float ub_y_coef;       // UB - LB grows by ub_y_coef * (staged energy): fp16 floor (+ the embedded scan's 2 x 16u)
    float thr_widen;       // thresholds read from the query state are widened by this factor (embedded scan)
    const float4 *tw2;     // 1024-point flavour: twiddle table, energy groups per piece, partial maxima per query
    int ncy, nqmax;
    int seed_group;        // 1024-point flavour: lanes per seed entry (32 / entries per warp)
    unsigned int stagger_ns;   // CTAs of the second half of the grid start this much later (see the kernel)
    unsigned long long *dbg;   // optional per-CTA timeline (8 globaltimer stamps per CTA), NULL in production
};

__device__ __forceinline__ int fft_pair_of_slot(const FftScanParams &p, int slot) {
    const unsigned long long prod = (unsigned long long)slot * (unsigned long long)p.perm;
    const unsigned long long qq = __double2ull_rz(__ull2double_rz(prod) * p.inv_np);
    long long pair = (long long)(prod - qq * (unsigned long long)p.npairs);
    if (pair < 0) pair += p.npairs;
    else if (pair >= p.npairs) pair -= p.npairs;
    return (int)pair;
}

// ONE warp: the threshold the query's histogram certifies (k entries at or below a bin), merged into
// s_thr[b] and published for the other CTAs.  Counts only grow, so every value read is a valid lower
// bound of the number of windows in its bin whatever the interleaving with other CTAs' increments.
__device__ __forceinline__ void fft_refresh_threshold(const FftScanParams &p, int b, float q2, float *s_thr) {
    const int lane = threadIdx.x & 31;
    unsigned int *hq = p.hist + (size_t)b * HSTRIDE;
    const uint2 c2 = __ldcg(reinterpret_cast<const uint2 *>(hq + HB) + lane);
    const unsigned int csum = c2.x + c2.y;
    unsigned int incl = csum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int u = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += u;
    }
    const unsigned int excl = incl - csum;
    const unsigned int hit = __ballot_sync(FULL, excl < p.k && p.k <= incl);
    if (hit == 0u) return;   // fewer than k entries so far
    const int src = __ffs(hit) - 1;
    // the coarse bin holding the k-th entry and the entries below it
    const unsigned int ex_s = __shfl_sync(FULL, excl, src), cx_s = __shfl_sync(FULL, c2.x, src);
    const bool second = ex_s + cx_s < p.k;
    const int cb = 2 * src + (second ? 1 : 0);
    const unsigned int below = ex_s + (second ? cx_s : 0u);
    // its 128 fine bins: 4 per lane
    const uint4 f4 = __ldcg(reinterpret_cast<const uint4 *>(hq + cb * HFINE_PER_COARSE) + lane);
    const unsigned int fsum = f4.x + f4.y + f4.z + f4.w;
    unsigned int fincl = fsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int u = __shfl_up_sync(FULL, fincl, o);
        if (lane >= o) fincl += u;
    }
    const unsigned int fexcl = below + fincl - fsum;
    int fine = 4 * lane;
    bool found = false;
    {
        unsigned int cum = fexcl;
        const unsigned int hv[4] = {f4.x, f4.y, f4.z, f4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (!found && cum < p.k && p.k <= cum + hv[i]) { fine = 4 * lane + i; found = true; }
            cum += hv[i];
        }
    }
    const unsigned int fhit = __ballot_sync(FULL, found);
    // (the fine counters may lag the coarse one: then the coarse bin's own upper edge is what is certified)
    const int edge = fhit ? cb * HFINE_PER_COARSE + __shfl_sync(FULL, fine, __ffs(fhit) - 1)
                          : cb * HFINE_PER_COARSE + HFINE_PER_COARSE - 1;
    if (lane == 0) {
        const int eb = hist_base(q2) + edge + 1;     // upper edge of the bin (exclusive)
        if (eb > 0 && eb < (0x7f800000 >> 13)) {
            const float tn = __uint_as_float((unsigned int)eb << 13) * p.widen2;
            atomicMin(reinterpret_cast<unsigned int *>(&s_thr[b]), __float_as_uint(tn));   // positive floats order as uints
            atomicMin(hq + H_THR, __float_as_uint(tn));
        }
    }
}

// Seeding pass: the CTA's 256 entries (one bin per thread, `cnt` false: none) are first counted in a
// shared-memory histogram (`sh`: HB uints, the idle exchange buffer) and only the non-empty bins go to the
// global histogram -- the per-thread minima of all CTAs fall into a few dozen fine bins and one or two
// coarse ones, and ~10^5 atomics on the same few addresses serialise in L2 (measured: 50 us per launch).
__device__ __forceinline__ void seed_hist_flush(unsigned int *hq, unsigned int *sh, int bin, bool cnt) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __syncthreads();                                   // every warp is done with the exchange buffers
#pragma unroll
    for (int i = 0; i < HB / 4 / fx2::THREADS; ++i) reinterpret_cast<uint4 *>(sh)[i * fx2::THREADS + tid] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    if (cnt) atomicAdd(sh + bin, 1u);
    __syncthreads();
    // warp w owns the bins [1024 w, 1024 w + 1024) = the coarse bins 8 w .. 8 w + 7
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        unsigned int csum = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int fb = 1024 * warp + 128 * m + 32 * j + lane;
            const unsigned int h = sh[fb];
            if (h != 0u) atomicAdd(hq + fb, h);
            csum += h;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) csum += __shfl_xor_sync(FULL, csum, o);
        if (lane == 0 && csum != 0u) atomicAdd(hq + HB + 8 * warp + m, csum);
    }
    __syncthreads();                                   // the exchange buffers may be overwritten
}

// The rare part of the epilogue (some window of the warp passed): per-window test, candidate append, upper
// bounds into the threshold histogram.  (Out of line it cost 17 %: the transform's registers went through the stack.)
__device__ __forceinline__ void fft_append_candidates(const FftScanParams &p, int b, const float2 (&v)[16], const __half2 *Ys,
                                                   bool any, float m2, float cu, float rhs, float inv_es, float base0,
                                                   float slack, int pair, int kb, bool count_ub) {
    const int tid = threadIdx.x, lane = tid & 31;
    const float INF = __int_as_float(0x7f800000);
    unsigned int *hq = p.hist + (size_t)b * HSTRIDE;
    const int hb = hist_base(p.st[b].q2);
    unsigned int mask = 0;
    if (any) {
#pragma unroll
        for (int d = 0; d < 16; ++d) {
            const float2 yf = __half22float2(Ys[tid + 256 * d]);
            const float2 val = fx2::fma2(v[d], make_float2(m2, m2), yf);
            if (val.x <= rhs && yf.x < INF) mask |= 1u << d;
            if (val.y <= rhs && yf.y < INF) mask |= 1u << (16 + d);
        }
    }
    const int cnt = __popc(mask);
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
    const long long ra = 2 * (long long)pair, rb = ra + 1;   // virtual rows
    const unsigned int fa = (unsigned int)((unsigned long long)(ra / p.nsegv) * (unsigned long long)p.Tp
                                           + (unsigned long long)(ra % p.nsegv) * (unsigned long long)p.hop);
    const unsigned int fb = (unsigned int)((unsigned long long)(rb / p.nsegv) * (unsigned long long)p.Tp
                                           + (unsigned long long)(rb % p.nsegv) * (unsigned long long)p.hop);
#pragma unroll
    for (int d = 0; d < 16; ++d) {
        if (mask & (0x10001u << d)) {
            const unsigned int t = (unsigned int)(kb + 256 * d);
            const float2 yf = __half22float2(Ys[tid + 256 * d]);
            const float2 ub2 = fx2::fma2(yf, make_float2(cu, cu), fx2::fma2(v[d], make_float2(m2, m2), yf));
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (mask & (1u << (16 * h + d))) {
                    if (pos < p.cap) dst[pos] = (h ? fb : fa) + t;
                    ++pos;
                    if (count_ub) {   // upper bound of the window's exact squared distance
                        const float u0 = h ? ub2.y : ub2.x;
                        const float ub = fmaxf((((u0 + 5.9604644775390625e-8f) * inv_es + base0) + 2.0f * slack) * 1.000001f, 0.0f);
                        const int bin = (int)(__float_as_uint(ub) >> 13) - hb;
                        if (ub < INF && bin < HB) hist_add_ub(hq, bin < 0 ? 0 : bin, 1u);
                    }
                }
            }
        }
    }
}

// End of a CTA's seeding pass: arrive (every histogram increment of this CTA is ordered before the arrival:
// barrier + fence); the CTA whose arrival is the `seed_need`-th derives the thresholds of all queries from
// the histograms and publishes them, every other CTA waits for that (bounded: arrivals never wait for
// anybody) and picks them up -- 300 CTAs re-deriving the same thresholds from the same cache lines took 3 us.
__device__ __forceinline__ void fft_seed_rendezvous(const FftScanParams &p, const float *s_q2, float *s_thr, int *s_flag) {
    const int tid = threadIdx.x;
    volatile unsigned int *done = p.hist + H_DONE;
    if (tid == 0) {
        __threadfence();
        const unsigned int ticket = atomicAdd(p.hist + H_ARR, 1u);
        *s_flag = (ticket + 1u == p.seed_need) ? 1 : 0;
        if (ticket + 1u != p.seed_need) {
            const unsigned long long t0 = globaltimer_ns();
            while (*done == 0u) {
                if (globaltimer_ns() - t0 > 2000000ull) break;   // 2 ms: thresholds stay loose, the call re-runs safely
                __nanosleep(100);
            }
        }
    }
    __syncthreads();
    if (*s_flag) {
        if (tid < 32) {
            for (int b = 0; b < p.nq; ++b) fft_refresh_threshold(p, b, s_q2[b], s_thr);
            __threadfence();
            if (tid == 0) *done = 1u;
        }
    } else if (tid < p.nq) {
        const unsigned int tb = *reinterpret_cast<volatile unsigned int *>(p.hist + (size_t)tid * HSTRIDE + H_THR);
        atomicMin(reinterpret_cast<unsigned int *>(&s_thr[tid]), tb);
    }
    __syncthreads();
}

// One CTA (256 threads) per row pair and iteration: Z^ * conj(Q)/N -> inverse FFT -> (D_a[t], D_b[t]);
// lower bound
//   LB = Q2 + Y2^ - 2 D^ - slack,
//   slack = 2 cf_u Qmax ynorm + 2 zqerr ||g|| + 12u (Q2 + ynorm^2)
// (FFT error; quantisation of the spectrum; Y2, Q2 roundings and the combination), tested in the pair's
// scaled units as fma(m2, v, yf) <= rhs.  Staging: the pair's spectrum (16 KiB) and its interleaved energy
// rows (16 KiB) + its statistics (16 bytes) arrive by TMA bulk copies issued one pair ahead.  With a single
// query its spectrum stays in registers (16 values per thread, a function of tid only).
//
// Pairs are handed out dynamically (an atomic slot counter; CTAs progress at different speeds), in a
// permuted order so that every prefix is a spread-out sample of the ensemble.
//
// Thresholds.  Every kept window adds its UPPER bound UB = LB + 2 slack + (fp16 floor of Y2) to the
// query's histogram; CTAs take turns re-deriving the threshold from it (one warp, no barrier) and publish
// it, every CTA picks the published value up once per pair.  p.seed: thresholds start at +inf; every CTA
// first evaluates its first pair WITHOUT appending anything -- each thread adds the minimum UB of its 32
// windows (256 distinct windows per CTA) -- and waits until `seed_need` CTAs have done so (a fraction of the
// grid: no co-residency assumption); the `seed_need`-th derives the first threshold and publishes it, and
// only then does a CTA test the pair's windows
// (one query: the transform's output is still in registers; a group of queries: the pair is transformed
// again).  This replaces round 1's separate seed launch.
//
// (Measured and rejected: ONE barrier per transform -- two alternating exchange buffers, the energy copy
// issued behind the barrier of its own pair -- 0.251 ms against 0.232 ms.)
//
// The small per-pair chores are spread over the warps (a warp that does all of them is late at every
// barrier): warp 0 issues the spectrum copy, warp 1 picks up published thresholds, warp 2 issues the
// energy copy, warp 3 draws the next pair, warp 4 re-derives thresholds.
template <bool SINGLEQ, bool EMB>
__global__ void __launch_bounds__(fx2::THREADS, 2) fft_scan_kernel(const FftScanParams p) {
    extern __shared__ __align__(128) unsigned char fsm[];
    __half2 *Zs = reinterpret_cast<__half2 *>(fsm);
    __half2 *Ys = Zs + fx2::N;
    float2 *ex = reinterpret_cast<float2 *>(Ys + fx2::N);
    float4 *pis = reinterpret_cast<float4 *>(ex + fx2::EX_FLOAT2);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(pis + 1);
    __shared__ float s_thr[QG_MAX], s_q2[QG_MAX], s_qmax[QG_MAX], s_gn[QG_MAX];
    __shared__ __align__(16) uint4 s_pub[QG_MAX];   // {published threshold bits, ...} of each query, one pair behind
    __shared__ int s_npair[2];                       // the pair behind the current one (-1: none), by pass parity
    __shared__ int s_flag;
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t barZ = smem_u32(&bars[0]), barY = smem_u32(&bars[1]);
    const float INF = __int_as_float(0x7f800000);

    const int slot = p.i0 + (int)blockIdx.x;
    if (slot >= p.i1) return;   // (a seeding launch is sized so that every CTA owns a pair)
#define PSH_STAMP(i) do { if (p.dbg != nullptr && tid == 0) p.dbg[(size_t)blockIdx.x * 8 + (i)] = globaltimer_ns(); } while (0)
    PSH_STAMP(0);
    // The two CTAs of an SM run the same code: started together they would sit in the same phase together --
    // both in a radix-16 pass (the packed fp32 pipe is saturated there), then both at a barrier or in an
    // exchange (it idles).  The second wave of CTAs starts half an iteration late so that the phases interleave.
    if (blockIdx.x >= (gridDim.x + 1) / 2 && p.stagger_ns != 0) {
        const unsigned long long t0 = globaltimer_ns();
        while (globaltimer_ns() - t0 < p.stagger_ns) __nanosleep(200);
    }
    if (tid == 0) {
        mbar_init(barZ, 1);
        mbar_init(barY, 1);
        mbar_fence_init();
    }
    if (tid < p.nq) {
        const float t0 = ld_volatile_f32(&p.st[tid].thr_fast);
        s_thr[tid] = EMB ? t0 * p.thr_widen : t0;
        s_q2[tid] = p.st[tid].q2;
        s_gn[tid] = p.st[tid].gnorm;
        s_qmax[tid] = 0.0f;
        s_pub[tid] = make_uint4(0x7f800000u, 0u, 0u, 0u);
    }
    __syncthreads();
    for (int b = 0; b < p.nq; ++b) {   // max_k |FFT(q)_k| from the QMAXP partial maxima (positive floats order as uints)
        static_assert(QMAXP == fx2::THREADS, "one partial maximum per thread");
        float m = __ldg(p.qmaxp + (size_t)b * QMAXP + tid);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
        if (lane == 0) atomicMax(reinterpret_cast<unsigned int *>(&s_qmax[b]), __float_as_uint(m));
    }
    __syncthreads();

    auto issue_z = [&](int pair) {  // one thread
        mbar_expect_tx(barZ, (uint32_t)(sizeof(__half2) * fx2::N));
        bulk_g2s(smem_u32(Zs), p.Z + (size_t)pair * fx2::N, (uint32_t)(sizeof(__half2) * fx2::N), barZ);
    };
    auto issue_y = [&](int pair) {  // one thread: the energy rows and the pair's statistics
        mbar_expect_tx(barY, (uint32_t)(sizeof(__half2) * fx2::N + sizeof(float4)));
        bulk_g2s(smem_u32(Ys), p.Y2 + (size_t)pair * fx2::N, (uint32_t)(sizeof(__half2) * fx2::N), barY);
        bulk_g2s(smem_u32(pis), p.pinfo + pair, (uint32_t)sizeof(float4), barY);
    };

    int pair = fft_pair_of_slot(p, slot);
    if (tid == 0) { issue_z(pair); issue_y(pair); }

    float2 qreg[16];
    if (SINGLEQ) {
#pragma unroll
        for (int i = 0; i < 16; ++i) qreg[i] = __ldg(p.Qc + tid + 256 * i);
    }
    const fx2::Seeds seeds = fx2::load_seeds(p.tw, tid);
    const int kb = fx2::out_base(tid);   // v[d] belongs to window kb + 256 d of both rows of the pair
    uint32_t phZ = 0, phY = 0;
    bool seeding = p.seed != 0;
    bool staged = false;        // this pair's spectrum and energies are already in shared memory (a group's pass after seeding)
    const bool rerun = p.nq > 1;  // seeding a group of queries: the pair is transformed twice
    const float cu = p.ub_y_coef;
    int par = 0;
    for (int iter = 0;;) {
        // warp 3 draws the slot behind this pair now and turns it into a pair right in front of the
        // transform's barrier -- the atomic's round trip hides behind the first radix-16 pass
        const bool draw = !(seeding && rerun);
        unsigned int drawn = 0;
        if (draw && tid == 96) drawn = atomicAdd(p.hist + H_SLOT, 1u);
        // warp 1: pick up the thresholds other CTAs have published -- fetched by an asynchronous 16-byte
        // copy during the previous pair, so nothing waits for L2 here -- and fetch the next ones
        // (visible to the CTA behind the transform's barrier)
        if (tid >= 32 && tid < 32 + p.nq) {
            const int b = tid - 32;
            cp_async_wait_all();
            const unsigned int tb = s_pub[b].x;
            if (tb < __float_as_uint(s_thr[b])) atomicMin(reinterpret_cast<unsigned int *>(&s_thr[b]), tb);
            cp_async_16(smem_u32(&s_pub[b]), p.hist + (size_t)b * HSTRIDE + H_THR);
        }
        if (!staged) { mbar_wait(barZ, phZ); phZ ^= 1; }
        for (int b = 0; b < p.nq; ++b) {
            float2 v[16];
            if (SINGLEQ) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = fx2::cmul(__half22float2(Zs[tid + 256 * i]), qreg[i]);
            } else {
                const float2 *Qb = p.Qc + (size_t)b * fx2::N;
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __ldg(Qb + tid + 256 * i);   // (L1/L2-resident; lands in v)
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = fx2::cmul(__half22float2(Zs[tid + 256 * i]), v[i]);
            }
            // behind the transform's only CTA barrier every thread has consumed the staged spectrum:
            // the next pair's copy is issued there (last query of the group)
            const bool last_q = b == p.nq - 1;
            fx2::ifft4096(v, ex, tid, seeds, [&]() {
                if (b == 0 && tid == 96) {
                    const long long ns = (long long)p.i0 + (long long)gridDim.x + (long long)drawn;
                    s_npair[par] = (draw && ns < (long long)p.i1) ? fft_pair_of_slot(p, (int)ns) : -1;
                }
            }, [&]() {
                if (last_q && tid == 0 && s_npair[par] >= 0) issue_z(s_npair[par]);
            });
            if (b == 0 && !staged) { mbar_wait(barY, phY); phY ^= 1; }  // the pair's window energies have landed
            const float4 pi = *pis;
            const float yn = pi.x, zq = pi.y, es = pi.z, m2 = pi.w;
            const float inv_es = 1.0f / es;               // es is a power of two
            const float q2 = s_q2[b], qmax = s_qmax[b], gn = s_gn[b];
            float slack;
            if (EMB) slack = (2.0f * p.cf_u * qmax * yn + 2.0f * zq * gn + p.slack_coef * q2 + p.g_coef * gn * yn) * 1.0001f;
            else slack = (2.0f * p.cf_u * qmax * yn + 2.0f * zq * gn + 7.152557373046875e-7f * (q2 + yn * yn)) * 1.0001f;
            const float base0 = q2 - slack;   // LB = (Y2^ - 2 D^) + base0, kept iff LB <= thr
            unsigned int *hq = p.hist + (size_t)b * HSTRIDE;
            const int hb = hist_base(q2);
            if (seeding) {
                // minimum over the thread's windows of the UPPER bound (scaled units); windows beyond T' are +inf
                float mn = INF;
#pragma unroll
                for (int d = 0; d < 16; ++d) {
                    const float2 yf = __half22float2(Ys[tid + 256 * d]);
                    const float2 ub2 = fx2::fma2(yf, make_float2(cu, cu), fx2::fma2(v[d], make_float2(m2, m2), yf));
                    mn = fminf(mn, fminf(ub2.x, ub2.y));
                }
                // true units; 2^-24: an energy in fp16's subnormal range was floored by at most that much
                const float ub = fmaxf((((mn + 5.9604644775390625e-8f) * inv_es + base0) + 2.0f * slack) * 1.000001f, 0.0f);
                int bin = (int)(__float_as_uint(ub) >> 13) - hb;
                bin = bin < 0 ? 0 : bin;
                const bool cnt = ub < INF && bin < HB;   // false for +inf (no valid window) and NaN
                seed_hist_flush(hq, reinterpret_cast<unsigned int *>(ex), cnt ? bin : 0, cnt);
                if (rerun) continue;
                PSH_STAMP(1);
                // one query: every increment of this CTA is ordered before its arrival (barrier + fence); wait
                // for `seed_need` arrivals (bounded: a CTA that runs arrives without waiting for anybody), derive
                // the threshold and test the windows whose transform is still in registers
                fft_seed_rendezvous(p, s_q2, s_thr, &s_flag);
                PSH_STAMP(3);
            }
            const float thr = s_thr[b];
            const float tdiff = thr - base0;
            const float rhs = fmaf(fabsf(tdiff), 9.5367431640625e-7f, tdiff) * es;   // (+2^-20: roundings of rhs and of the fma below)
            // v[d] = (D_a, D_b) of window kb + 256 d; the energies are stored in that order, both rows of the
            // pair in one word: one FFMA2 per pair of windows
            float mn = INF;
#pragma unroll
            for (int d = 0; d < 16; ++d) {
                const float2 val = fx2::fma2(v[d], make_float2(m2, m2), __half22float2(Ys[tid + 256 * d]));
                mn = fminf(mn, fminf(val.x, val.y));
            }
            const bool any = mn <= rhs;   // (+inf <= +inf while thr = +inf: sorted out per window below)
            if (__any_sync(FULL, any))   // rare
                fft_append_candidates(p, b, v, Ys, any, m2, cu, rhs, inv_es, base0, slack, pair, kb, !(p.seed != 0 && iter == 0));
            __syncthreads();  // the exchange buffer (and, after the last query, the energy rows) may be overwritten
        }
        if (seeding && rerun) {
            // a group of queries: arrive, wait for the thresholds, then the same pair again
            fft_seed_rendezvous(p, s_q2, s_thr, &s_flag);
            seeding = false;
            staged = true;
            par ^= 1;
            continue;
        }
        seeding = false;
        staged = false;
        const int npair = s_npair[par];                   // (the next pass writes the other slot)
        par ^= 1;
        if (tid == 64 && npair >= 0) issue_y(npair);      // warp 2
        if (iter == 0) PSH_STAMP(4);
        if (iter == 1) PSH_STAMP(5);
        if (iter == 8) PSH_STAMP(6);
        if (npair < 0) { PSH_STAMP(7); break; }
        if ((tid >> 5) == 4 && ((iter + (int)blockIdx.x) & p.refresh_mask) == 0)   // warp 4
            for (int b = 0; b < p.nq; ++b) fft_refresh_threshold(p, b, s_q2[b], s_thr);
        pair = npair;
        ++iter;
    }
}
