/*
 * pshadow.h -- C ABI of libpshadow.so: the B200 (sm_100a) path-shadowing scan.
 *
 * The reference (RudyMorel/shadowing @ 751a800) is pure Python and has no FFI of its own; its
 * plugin surface is the Python classes in shadowing/path_shadowing/*.py.  This header is the
 * boundary a binding for that path would target: every entry point names the reference code it
 * replaces.  Plain pointers and sizes only; all pointers prefixed d_ are DEVICE pointers owned
 * by the caller; all work is enqueued on the caller's cudaStream_t (passed as void*).
 *
 * Return value of every int function: 0 = ok, < 0 = PSH_E_* argument/state error,
 * > 0 = a cudaError_t raised by the runtime.  Nothing here throws, aborts or prints.
 */
#ifndef PSHADOW_H
#define PSHADOW_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSH_VERSION 100 /* 0.1.0 */

enum {
    PSH_OK = 0,
    PSH_E_ARG = -1,        /* null pointer / non-positive size / W+H > T                      */
    PSH_E_K = -2,          /* k exceeds the number of windows (reference: torch.topk raises)  */
    PSH_E_WORKSPACE = -3,  /* workspace smaller than psh_scan_workspace_bytes()               */
    PSH_E_TOO_LARGE = -4,  /* R*T' >= 2^32 windows in one call: shard the rows and merge      */
    PSH_E_UNSUPPORTED = -5, /* context too long for the kernel asked for (direct scans: not   */
                            /* even one query's staging fits shared memory; fft: W > 2048)    */
    PSH_E_OVERFLOW = -6     /* psh_scan_overflowed: repeat the scan without PSH_FLAG_NOSYNC   */
};

/* scan modes */
enum {
    PSH_MODE_EXACT = 0, /* every window evaluated with the reference's exact fp32 sequence    */
    PSH_MODE_FILTER = 1, /* 1-FMA/element lower-bound filter, exact re-rank of the survivors; */
                         /* results identical to PSH_MODE_EXACT by construction               */
    PSH_MODE_FFT = 2     /* lower-bound filter through one inverse FFT per trajectory pair    */
                         /* (needs psh_fft_prepare aux, else behaves as FILTER; trajectories  */
                         /* longer than 4096 samples are cut into overlapping pieces);        */
                         /* same exact re-rank, results identical to PSH_MODE_EXACT           */
};

/* OR-ed into `mode`: enqueue only, do not synchronise; the caller must call
 * psh_scan_overflowed() (which synchronises) before trusting the results. */
#define PSH_FLAG_NOSYNC 0x100
/* OR-ed into `mode`: this scan is one of a pipeline alternating between streams -- its persistent scan kernel
 * leaves a few SMs to the other streams' small kernels (re-rank, select, exchange), which would otherwise
 * wait until the scan has drained.  PSH_SHARE_SMS(n) leaves n SMs (n = 1..63); the bare flag leaves six
 * (PSH_SPARE_SMS overrides both).  Results are unaffected. */
#define PSH_FLAG_SHARE_SMS 0x200
#define PSH_SHARE_SMS(n) (PSH_FLAG_SHARE_SMS | (((n) & 0x3f) << 12))

int psh_version(void);
const char *psh_error_string(int code);

/* Bytes of device scratch psh_scan_topk_f32 needs for these sizes (0 on invalid sizes). */
size_t psh_scan_workspace_bytes(int64_t R, int64_t T, int B, int W, int H, int64_t k);

/*
 * The scan: replaces PathShadowing.batched_distance (path_shadowing.py:97-179) for
 * Identity embedding (path_embedding.py:117-139 + pad_context :48-51) and RelativeMSE
 * (path_distance.py:62-65), including the per-split torch.topk + running merge (:165-173)
 * and select_cartesian_product (:43-58).
 *
 *   d_dataset   (R rows, row_stride floats apart, T valid samples each), fp32
 *   d_queries   (B, W) contiguous fp32 contexts
 *   H           PredictionContext horizon (0 for None): windows t = 0 .. T-W-H
 *   k           neighbours kept per query
 *   row_offset  added to the returned trajectory index (rank's first global row when sharded)
 *   d_out_dist  (B, k) fp32, ascending
 *   d_out_idx   (B, k, 2) int32 [trajectory, offset]; ties ordered by (distance, r*T'+t).
 *               NULL: d_out_dist receives packed (B, k, 3) int32 records [distance bits,
 *               trajectory, offset] instead (the payload of the multi-GPU all-gather)
 *   d_ws        scratch of >= psh_scan_workspace_bytes(...) bytes, 256-byte aligned
 *   d_aux       PSH_MODE_FFT: buffer filled by psh_fft_prepare for this dataset/W/H, else NULL
 *
 * Distances carry the reference's CPU bit pattern: s = sum_j fl(fl(q_j - y_{t+j})^2)
 * accumulated sequentially in fp32 without FMA, sqrt, IEEE divide by ||q|| (8-lane order).
 * Synchronises the stream once before returning (overflow check of the candidate buffers).
 */
int psh_scan_topk_f32(const float *d_dataset, int64_t R, int64_t T, int64_t row_stride,
                      const float *d_queries, int B, int W, int H, int64_t k,
                      int32_t row_offset, int mode,
                      float *d_out_dist, int32_t *d_out_idx,
                      void *d_ws, size_t ws_bytes, const void *d_aux, size_t aux_bytes, void *stream);

/*
 * The scan in EMBEDDED space: replaces PathShadowing.batched_distance (path_shadowing.py:97-179)
 * for any LINEAR embedding -- PathEmbedding(kernel) with a (d, 1, W) kernel, e.g. Foveal
 * (path_embedding.py:117-132, 142-172) -- and RelativeMSE over the d embedded dimensions
 * (path_distance.py:62-65).  The reference embeds every window with conv1d (kernel zero-padded by
 * H); here the kernel arrives decomposed into runs of equal taps and every window is evaluated
 * from a prefix sum of its staged row segment:  e_n(t) = sum_runs c (P[t+b] - P[t+a]).
 *
 *   d_qemb   (B, d) contiguous fp32: the EMBEDDED contexts (the caller embeds the few query
 *            windows itself, as path_shadowing.py:138 does)
 *   d_runs   (nruns) records {int32 row, int32 a, int32 b, float c}: kernel[row][a..b) == c,
 *            0 <= a < b <= W, rows ascending (rows without runs are all-zero taps)
 *   flags    0, PSH_FLAG_NOSYNC and / or PSH_FLAG_SHARE_SMS;   everything else as psh_scan_topk_f32 (same workspace size)
 *
 * Distances follow the reference's sequence over the embedded dimensions (s accumulated n
 * ascending, non-fused; sqrt; divide by ||ex|| in torch's order); the embedded values themselves
 * are box sums accurate to ~2 ulp, whereas the reference's conv1d rounds in an order that cannot be
 * replayed: parity with the reference is within 1e-6 relative, indices up to near-ties.
 */
int psh_scan_topk_embed_f32(const float *d_dataset, int64_t R, int64_t T, int64_t row_stride,
                            const float *d_qemb, int B, int d, int W, int H, int64_t k,
                            int32_t row_offset, int flags, const void *d_runs, int nruns,
                            const float *d_g, const void *d_aux, size_t aux_bytes,
                            float *d_out_dist, int32_t *d_out_idx, void *d_ws, size_t ws_bytes, void *stream);
/*
 * FFT flavour of the embedded scan (d_aux != NULL): ||ex - K y_t||^2 = ||ex||^2 - 2 g.y_t + ||K y_t||^2
 * with g = K^T ex, so the cross term is one correlation per trajectory -- the Identity flavour's
 * spectra and inverse FFT -- and the quadratic term is precomputed per dataset and kernel:
 *   d_g     (B, W) fp32: g_b = K^T ex_b (the caller multiplies; B x d x W flops)
 *   d_aux   buffer of psh_fft_aux_bytes(R, T, W, H) bytes filled by psh_fft_prepare_embed for this
 *           dataset, kernel (runs), W and H.  NULL: every window is evaluated exactly.
 * The filter is a rigorous lower bound; survivors are re-evaluated with the exact embedded
 * arithmetic, so both flavours return the same windows (distances equal up to the last bit of
 * a box sum).
 */
int psh_fft_prepare_embed(const float *d_dataset, int64_t R, int64_t T, int64_t row_stride, int W, int H,
                          const void *d_runs, int nruns, void *d_aux, size_t aux_bytes, void *stream);

/* After one or more PSH_FLAG_NOSYNC scans on the same workspace (and whatever the caller enqueued
 * behind them): synchronise the stream and report PSH_OK, or PSH_E_OVERFLOW if a candidate buffer
 * overflowed in ANY of those scans since the previous check (adversarially ordered data; the
 * flag is sticky in the workspace and cleared by this call) -- then their outputs are invalid
 * and they must be repeated without the flag. */
int psh_scan_overflowed(const void *d_ws, int B, void *stream);

/*
 * Dataset-side precomputation for PSH_MODE_FFT (no reference counterpart; the reference
 * recomputes everything per call).  Fills d_aux (>= psh_fft_aux_bytes, 256-byte aligned) with
 * the 4096-point spectra of all trajectory pairs quantised to fp16 pairs (per-pair power-of-two
 * scale, measured quantisation error), the window energies sum_{j<W} y_{t+j}^2 scaled and rounded
 * DOWN to fp16, and per-pair statistics: 8 bytes per pair of samples, the size of the raw rows.
 * Valid for this (dataset, W, H) only; W <= 2048 (else PSH_E_UNSUPPORTED / 0); a trajectory longer
 * than one 4096-point transform is cut into overlapping 4096-sample pieces (overlap-save).
 */
size_t psh_fft_aux_bytes(int64_t R, int64_t T, int W, int H);
int psh_fft_prepare(const float *d_dataset, int64_t R, int64_t T, int64_t row_stride, int W, int H,
                    void *d_aux, size_t aux_bytes, void *stream);

/* Test hook: n independent 4096-point complex transforms (dir -1 forward, +1 inverse, unnormalised,
 * table twiddles; dir 3: the scan's own packed-fp32 inverse) with the library's FFT; d_aux is a
 * prepared aux buffer (twiddles). */
int psh_debug_fft4096(const void *d_in, void *d_out, int n, int dir, const void *d_aux, void *stream);
/* Test hook: n independent 1024-point complex transforms (dir +1: the scan's warp-level inverse, unnormalised;
 * dir -1: the forward transform psh_fft_prepare uses for 1024-sample pieces), natural order in and out;
 * d_aux is any prepared aux buffer (twiddles). */
int psh_debug_fft1024(const void *d_in, void *d_out, int n, int dir, const void *d_aux, void *stream);

/*
 * k-way merge of G per-shard results into the global top-k: replaces the cat + topk +
 * fancy-index merge (path_shadowing.py:170-173) across GPUs.
 *   d_dist_parts (G, B, k) fp32, d_idx_parts (G, B, k, 2) int32 (global trajectory ids); every
 *                shard's k records ascending in (distance, r*Tp+t), as psh_scan_topk_f32 writes them
 *   Tp           windows per trajectory (tie order is (distance, r*Tp+t))
 */
int psh_merge_topk(const float *d_dist_parts, const int32_t *d_idx_parts, int G, int B,
                   int64_t k, int64_t Tp, float *d_out_dist, int32_t *d_out_idx, void *stream);
/* Same merge on packed records (G, B, k, 3) int32 = [distance bits, trajectory, offset]: the
 * layout one ncclAllGather of the per-rank results produces. */
int psh_merge_topk_packed(const int32_t *d_rec_parts, int G, int B, int64_t k, int64_t Tp,
                          float *d_out_dist, int32_t *d_out_idx, int32_t *d_overflow_flag, void *stream);
/* d_overflow_flag (may be NULL; zero it first): set to 1 if any shard's PSH_FLAG_NOSYNC scan
 * overflowed a candidate buffer (it poisons its first record); every rank sees the same flag,
 * so all ranks repeat the step with synchronous scans together. */

/*
 * Multi-GPU, one process per GPU: all-gather of the per-rank records over NVLink peer memory fused
 * with the merge -- ONE kernel instead of ncclAllGather + psh_merge_topk_packed (no reference
 * counterpart; the reference is single-process).  Every rank creates an exchange buffer of
 * psh_xchg_bytes(G, B, k) bytes (cudaMalloc, zeroed), publishes its 64-byte CUDA IPC handle, and
 * maps the other ranks' buffers with psh_xchg_open.  psh_allgather_merge_packed is collective:
 * every rank calls it with the same (B, k, Tp) and the same epoch = 1, 2, 3, ... (one per call);
 *   d_rec_local (B, k, 3) int32 records of this rank (psh_scan_topk_f32 with d_out_idx = NULL)
 *   bufs        HOST array of G device pointers, bufs[g] = rank g's exchange buffer as mapped in
 *               this process (bufs[rank] = the own buffer)
 *   d_flag      int32, zeroed by the caller: bit 0 = a shard's NOSYNC scan overflowed,
 *               bit 1 = a peer's records did not arrive within 30 s (PSH_XCHG_TIMEOUT_MS; results invalid)
 * G CTAs per query: CTA (g, b) stores query b's records into rank g's buffer -- every datum in one
 * 8-byte store together with the epoch, so the words validate themselves: no fence, no flag (fallback
 * for G*k*12 bytes > 200 KB: plain stores, a system-scope fence and per-(rank, query) flags) --, polls
 * the G record sets of its query and places the records of list g (same order as psh_merge_topk).
 */
size_t psh_xchg_bytes(int G, int B, int64_t k);
int psh_xchg_create(size_t bytes, void **d_buf, unsigned char *ipc_handle_64);
int psh_xchg_open(const unsigned char *ipc_handle_64, void **d_peer);
int psh_xchg_close(void *d_peer);
int psh_xchg_destroy(void *d_buf);
int psh_allgather_merge_packed(const int32_t *d_rec_local, void *const *bufs, int G, int rank, int B, int64_t k,
                               int64_t Tp, uint32_t epoch, float *d_out_dist, int32_t *d_out_idx,
                               int32_t *d_flag, void *stream);
/* The same step in two launches -- psh_xchg_send (stores + flags, never waits) and psh_xchg_merge
 * (wait for the G flags of the epoch, merge) -- so a pipeline of scans can enqueue
 *     scan(i+1), send(i+1), merge(i)
 * and a rank computes its next scan instead of idling until the slowest peer has delivered step i.
 * The exchange buffers hold eight epochs (epoch % 8): two per stream of a pipeline that alternates steps
 * between up to four streams (this order needs two as well). */
int psh_xchg_send(const int32_t *d_rec_local, void *const *bufs, int G, int rank, int B, int64_t k,
                  uint32_t epoch, void *stream);
int psh_xchg_merge(void *const *bufs, int G, int rank, int B, int64_t k, int64_t Tp, uint32_t epoch,
                   float *d_out_dist, int32_t *d_out_idx, int32_t *d_flag, void *stream);

/*
 * Gather the winning paths with their out-context: replaces path_shadowing.py:210-216.
 *   d_idx (n, 2) int32 [trajectory - row_offset must be in [0,R)], d_out (n, L) fp32, L = W+H.
 *   Rows whose trajectory falls outside [row_offset, row_offset+R) are written as zeros
 *   (sharded gather: the owner writes, a sum over ranks assembles).
 */
int psh_gather_paths(const float *d_dataset, int64_t R, int64_t T, int64_t row_stride,
                     const int32_t *d_idx, int64_t n, int32_t row_offset, int L,
                     float *d_out, void *stream);

/*
 * Fused realised-variance + weighted aggregation: replaces predict_from_paths
 * (path_shadowing.py:234-254) for to_predict = realized_variance(., Ts, vol)[:, :, 0, :]
 * (statistics.py:5-16) and proba Uniform / Softmax (scatspectra; w ~ exp(-d^2 / 2 eta^2)).
 *   d_paths (B, k, L) fp32, the last H samples are the out-context; d_dist (B, k)
 *   d_Ts (nT) int32 maturities (<= H);  proba: 0 uniform, 1 softmax;  vol: 0/1
 *   d_mean, d_std (B, nT) fp32
 */
int psh_rv_aggregate(const float *d_paths, const float *d_dist, int B, int64_t k, int L, int H,
                     const int32_t *d_Ts, int nT, float eta, int proba, int vol,
                     float *d_mean, float *d_std, void *stream);

/* Number of kernels libpshadow has launched in this process (bench.py's gpu_launches). */
uint64_t psh_launch_count(void);

/*
 * Measurement hooks (no reference counterpart): between begin and end every kernel the library
 * launches is bracketed by CUDA events on the caller's stream.  psh_profile_end waits for them
 * and returns summed milliseconds and launch counts per kind: 0 = scan kernels, 1 = select /
 * re-rank / finalise kernels, 2 = merge / peer-memory all-gather + merge kernels.  State is per
 * calling thread (begin, the launches and end must come from the same thread); bench.py's roofline leg.
 */
void psh_profile_begin(void);
int psh_profile_end(double *ms_by_kind, uint64_t *launches_by_kind, int nkinds);

#ifdef __cplusplus
}
#endif
#endif /* PSHADOW_H */
