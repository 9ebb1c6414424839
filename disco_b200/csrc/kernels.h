// Internal launcher interface between the kernel translation units and the C ABI (api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace disco {

// SM count of the current device (cached per device; api.cu) -- launch heuristics size their grids with it
int sm_count();
// grid.y / grid.z are limited to 65535 blocks
constexpr int kMaxGridYZ = 65535;

struct StftArgs {
    const float* x;         // [n_sig][L] float32 time signals (n_sig = n_grp * C, last group may be short)
    const float* mask;      // SCM only: [n_grp][T][F] (mask_ft = 0) or [n_grp][F][T] (mask_ft = 1)
    const float* mask2;     // two-mask SCM only: the second mask, same layout
    float2* Y;              // [n_sig][T][F] complex64, frame-major
    float* part;            // SCM only: [n_grp][slots_per_grp][n_mask][2 C^2][F] partial sums per (group, CTA) segment
    const float2* twiddle;  // [N/32][32]: W_N^(l*k1)
    const float* window;    // [N]: 0.5 * periodic Hann
    int n_sig, n_grp, L, T;
    int slots_per_grp;
    int mask_ft;
    int use_tma;
};

// n_mask: 0 plain STFT, 1 STFT + SCMs under `mask`, 2 STFT + SCMs under `mask` and `mask2`
cudaError_t launch_stft_scm(const StftArgs& a, int n_fft, int C, int n_cta, int n_mask, cudaStream_t st);
bool stft_scm_supported(int n_fft, int C, int n_mask);
// matrices of mask set `set` (of n_set) from the segment partial sums
cudaError_t launch_scm_finalize(const float* part, float2* Rss, float2* Rnn, int n_grp, int slots_per_grp,
                                int tiles_per_grp, int n_cta, int C, int F, int T, int n_set, int set, cudaStream_t st);
int stft_tile_frames(int n_fft, int C);
int stft_tiles_per_grp(int n_fft, int C, int T);
int stft_slots_per_grp(int n_grp, int tiles_per_grp, int n_cta);
int stft_cta_of_tile_host(long long i, long long total, int nb);

// Step-2 style input: group g = (utterance b, node k) sees D = C + K - 1 channels:
// its own C microphone spectra, then the compressed signals z of the other nodes in node
// order (reference concatenate_signals, tango.py:142-155).
// Ragged arrays (nodes with different microphone counts) are handled by launching once per
// channel count on the subset `sel` of nodes that have C microphones: group g = (b, sel[g % n_sel]).
struct CatArgs {
    const float2* Y;   // [n_grp][C][T][F]
    const float2* Z;   // [n_utt][K][T][F] (z_sb = K, z_sk = 1) or node-major [K][n_utt][T][F] (z_sb = 1, z_sk = n_utt);
                       // may be null when K == 1
    long long z_sb, z_sk;   // plane (T x F) strides of Z along the utterance and the node axis
    int C, K, T, F;
    int n_grp;         // = n_utt * n_sel
    int n_sel;         // nodes covered by this launch (K when all nodes have C microphones)
    int sel[16];       // their node indices, ascending
};

struct ScmArgs {
    CatArgs in;
    const float* mask;   // [n_grp][T][F] or [n_grp][F][T]; null = all ones (plain SCM into Rss, Rnn = 0)
    int mask_ft;
    float2* Rss;         // [n_grp][F][D][D]
    float2* Rnn;
    // optional fused step-1 filter-and-sum (single-node groups, K == 1): z = w1^H y, zn = y[ref] - z
    const float2* W1;    // [n_grp][F][C] or null
    float2* z_out;       // [n_grp][T][F]
    float2* zn_out;      // [n_grp][T][F] or null
    int ref;
};
cudaError_t launch_masked_scm(const ScmArgs& a, cudaStream_t st);
cudaError_t launch_masked_scm_wide(const ScmArgs& a, cudaStream_t st);   // D = 5..16 (scm_wide.cu)

struct SolveArgs {
    const float2* Rss;   // [n_mat][D][D]
    const float2* Rnn;
    float2* W;           // [n_mat][D]
    float2* T1;          // [n_mat][D] (may be null)
    int n_mat, D;
    int type;            // 0 gevd, 1 r1-mwf, 2 mwf
    int rank;            // gevd: number of generalised eigenpairs kept; <= 0 or >= D means full
    double mu;
    // optional: read the matrices straight from the fused STFT+SCM kernel's segment partial sums
    // (skips scm_finalize); matrix idx = (set * n_grp + grp) * F + f, n_mat = n_set * n_grp * F
    const float* part;   // [n_grp][slots_per_grp][n_set][2 D^2][F] or null
    int slots_per_grp, tiles_per_grp, n_cta, F;
    int n_set;           // mask sets in the partial sums (0 is read as 1)
    float inv_T;
};
cudaError_t launch_mwf_solve(const SolveArgs& a, cudaStream_t st);

struct FilterArgs {
    CatArgs in;
    const float2* W;     // [n_grp][F][D]
    int conj_w;          // 1: w^H x (reference np.inner(conj(w), x)); 0: w^T x (reference np.inner(t1, x))
    float2* out;         // [n_grp][T][F] (out_ft = 0) or [n_grp][F][T] (out_ft = 1)
    float2* resid;       // optional: in[ref] - out, same layout as out (reference zn, tango.py:376)
    int ref;             // reference channel for resid
    int out_ft;
};
cudaError_t launch_filter_sum(const FilterArgs& a, cudaStream_t st);
cudaError_t launch_filter_sum_multi(const FilterArgs& a, cudaStream_t st);   // K > 1, all nodes, TF output

// Single-node groups, both filters in one pass over Y (filter_dual.cu): z = w1^H y, zn = y[ref] - z, yf = w2^H y
struct DualFilterArgs {
    const float2* Y;     // [n_grp][C][T][F]
    const float2* W1;    // [n_grp][F][C] step-1 filters
    const float2* W2;    // [n_grp][F][C] step-2 filters
    float2* z;           // [n_grp][T][F] (out_ft = 0) or [n_grp][F][T]
    float2* zn;          // same layout, may be null
    float2* yf;          // same layout
    int n_grp, C, T, F, ref, out_ft;
};
cudaError_t launch_filter_dual(const DualFilterArgs& a, int sm_count, cudaStream_t st);

// Fused multi-node middle pass (mid_multi.cu): z, zn of every node + step-2 SCMs of every node.
struct MidArgs {
    const float2* Y;     // [B*K][C][T][F]
    const float2* W1;    // [B*K][F][C]  step-1 filters
    const float* mask;   // [B*K][T][F]  step-2 masks (frame-major only)
    float2* Z;           // [B][K][T][F] out
    float2* ZN;          // [B][K][T][F] out (may be null)
    float2* Rss;         // [B*K][F][D][D] out, D = C + K - 1, reference channel order
    float2* Rnn;
    int B, K, C, T, F, ref;
};
cudaError_t launch_tango_mid(const MidArgs& a, cudaStream_t st);
bool tango_mid_supported(int C, int K);

// Recursive (online) SCMs and block-wise filtering (online.cu; reference internal_formulas.py:84-103).
struct OnlineArgs {
    CatArgs in;
    const float* mask;      // [n_grp][T][F] frame-major, or null (all ones: plain smoothed SCM into Rss, Rnn = decay only)
    const float2* R0ss;     // optional initial matrices [n_grp][F][D][D]
    const float2* R0nn;
    float2* Rss;            // [n_grp][J][F][D][D]: smoothed SCMs after the last frame of every block
    float2* Rnn;
    int P, J;               // frames per block, number of blocks = ceil(T / P)
    int power;              // 2: weights m^2, (1-m)^2 (x = m y estimate, M = None); 1: m, 1-m (x = mixture, M = mask)
    float lam_block, lam_last;   // lambda^P, lambda^(frames of the last block)
    float gw[64];           // (1 - lambda) lambda^k, k = 0..P-1
};
cudaError_t launch_scm_recursive(const OnlineArgs& a, cudaStream_t st);

struct OnlineFilterArgs {
    CatArgs in;
    const float2* W;        // [n_grp][J][F][D] one filter per block
    int conj_w;
    float2* out;            // [n_grp][T][F]
    float2* resid;          // optional x[ref] - out
    int ref, P, J, lag;     // frame t uses filter t / P - lag (pass-through of channel `ref` while that is < 0)
};
cudaError_t launch_filter_sum_blocks(const OnlineFilterArgs& a, cudaStream_t st);

// IIR filter bank + band statistics (filterbank.cu; reference metrics.py fw_snr / fw_sd).
struct BankArgs {
    const float* x;      // [n_sig] rows of L samples, row stride ldx
    const float* sel;    // optional, same layout: statistics over samples with sel != 0 (else: output != 0)
    const double* ba;    // [n_band][2][order + 1]: numerator b, then denominator a
    double* stats;       // [n_sig][n_band][3]: count, sum, sum of squares of the selected filter outputs
    int n_sig, L, n_band;
    long long ldx;
};
cudaError_t launch_band_stats(const BankArgs& a, int order, cudaStream_t st);

struct IstftArgs {
    const float2* Y;     // [n_sig][T][F] frame-major complex64
    float* x;            // [n_sig][L]
    const float2* twiddle;
    const float* window; // [N] periodic Hann (unscaled)
    int n_sig, L, T;
};
cudaError_t launch_istft(const IstftArgs& a, int n_fft, cudaStream_t st);

cudaError_t launch_tf_mask(const float2* S, const float2* Nn, float* M, size_t n, int kind, int power,
                           float thr_lin, cudaStream_t st);
// out[b][c][r] = in[b][r][c]
cudaError_t launch_transpose_c64(const float2* in, float2* out, int batch, int rows, int cols, cudaStream_t st);
cudaError_t launch_transpose_f32(const float* in, float* out, int batch, int rows, int cols, cudaStream_t st);
// out = m * in or (1 - m) * in over n = n_grp * chans * plane points; the mask [n_grp][plane] is shared by the
// `chans` channels of a group
cudaError_t launch_apply_mask(const float2* in, const float* m, float2* out, size_t n, size_t plane, int chans,
                              int one_minus, cudaStream_t st);

}  // namespace disco
