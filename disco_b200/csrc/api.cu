// C ABI of libdisco_b200.so (declared in include/disco_b200.h): argument checking, per-device
// constant tables, launch-geometry choices.  No torch types cross this boundary.
#include <math.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/disco_b200.h"
#include "kernels.h"

using namespace disco;

namespace disco {
// SM count of the CURRENT device (cached per device: processes may drive several GPUs)
int sm_count() {
    static int cache[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cache[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cache[dev] = n;
    }
    return cache[dev];
}

}  // namespace disco

namespace {

thread_local std::string g_err;

int fail(int code, const char* msg) {
    g_err = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char* where) {
    g_err = std::string(where) + ": " + cudaGetErrorString(e);
    return (int)e;
}
#define CU(expr, where)                                  \
    do {                                                 \
        cudaError_t _e = (expr);                         \
        if (_e != cudaSuccess) return cuda_fail(_e, where); \
    } while (0)

bool valid_nfft(int n) { return n == 256 || n == 512 || n == 1024; }

struct Tables {
    float2* twiddle = nullptr;   // [N/32][32] W_N^(l k1)
    float* win_half = nullptr;   // 0.5 * periodic Hann (forward: two-for-one split needs the 1/2)
    float* win = nullptr;        // periodic Hann
};
std::mutex g_mu;
std::map<std::pair<int, int>, Tables> g_tables;  // (device, n_fft)

// Build (once per device and n_fft) the twiddle and window tables, computed in double on the host.
int get_tables(int n_fft, Tables* out) {
    int dev = 0;
    CU(cudaGetDevice(&dev), "cudaGetDevice");
    std::lock_guard<std::mutex> lk(g_mu);
    auto key = std::make_pair(dev, n_fft);
    auto it = g_tables.find(key);
    if (it != g_tables.end()) {
        *out = it->second;
        return 0;
    }
    const int RA = n_fft / 32;
    std::vector<float2> tw(n_fft);
    std::vector<float> wh(n_fft), w(n_fft);
    const double two_pi = 6.283185307179586476925286766559;
    for (int k1 = 0; k1 < RA; ++k1)
        for (int l = 0; l < 32; ++l) {
            const double ang = -two_pi * (double)(l * k1) / (double)n_fft;
            tw[k1 * 32 + l] = make_float2((float)cos(ang), (float)sin(ang));
        }
    for (int n = 0; n < n_fft; ++n) {
        const double h = 0.5 - 0.5 * cos(two_pi * (double)n / (double)n_fft);
        w[n] = (float)h;
        wh[n] = (float)(0.5 * h);
    }
    Tables t;
    CU(cudaMalloc(&t.twiddle, n_fft * sizeof(float2)), "cudaMalloc tables");
    CU(cudaMalloc(&t.win_half, n_fft * sizeof(float)), "cudaMalloc tables");
    CU(cudaMalloc(&t.win, n_fft * sizeof(float)), "cudaMalloc tables");
    CU(cudaMemcpy(t.twiddle, tw.data(), n_fft * sizeof(float2), cudaMemcpyHostToDevice), "cudaMemcpy tables");
    CU(cudaMemcpy(t.win_half, wh.data(), n_fft * sizeof(float), cudaMemcpyHostToDevice), "cudaMemcpy tables");
    CU(cudaMemcpy(t.win, w.data(), n_fft * sizeof(float), cudaMemcpyHostToDevice), "cudaMemcpy tables");
    g_tables[key] = t;
    *out = t;
    return 0;
}

// Persistent launch geometry of the fused STFT kernel: one CTA per SM (fewer when there are fewer tiles).
struct StftPlan {
    int n_cta, tiles_per_grp, slots_per_grp;
};
// SMs the persistent fused kernel leaves free (disco_set_reserved_sms): room for the CTAs of a concurrent NCCL
// collective, which cannot co-reside with a 213 KB-shared-memory CTA
int g_reserved_sms = 0;

StftPlan plan_stft(int n_grp, int C, int T, int n_fft) {
    StftPlan pl;
    pl.tiles_per_grp = stft_tiles_per_grp(n_fft, C, T);
    const long long total = (long long)n_grp * pl.tiles_per_grp;
    int sms = sm_count() - g_reserved_sms;
    if (sms < 1) sms = 1;
    pl.n_cta = (int)(total < sms ? total : sms);
    pl.slots_per_grp = stft_slots_per_grp(n_grp, pl.tiles_per_grp, pl.n_cta);
    return pl;
}

size_t stft_ws_bytes(int n_grp, int C, int length, int n_fft, int n_mask) {
    if (!valid_nfft(n_fft) || C < 1 || C > 8 || n_grp < 1 || n_mask < 1) return 0;
    const int T = disco_n_frames(length, n_fft);
    const StftPlan pl = plan_stft(n_grp, C, T, n_fft);
    return (size_t)n_grp * pl.slots_per_grp * n_mask * 2 * C * C * (n_fft / 2 + 1) * sizeof(float);
}

int stft_common(const float* x, const float* mask, const float* mask2, int mask_layout, void* Y, void* Rss,
                void* Rnn, int n_sig, int C, int length, int n_fft, void* workspace, size_t workspace_bytes, int n_mask,
                void* stream) {
    if (!valid_nfft(n_fft)) return fail(DISCO_ERR_INVALID, "n_fft must be 256, 512 or 1024");
    if (n_sig <= 0 || length <= n_fft / 2)
        return fail(DISCO_ERR_INVALID, "need n_sig > 0 and length > n_fft/2 (reflect padding)");
    if (C < 1) return fail(DISCO_ERR_INVALID, "C must be positive");
    if (!stft_scm_supported(n_fft, C, n_mask))
        return fail(DISCO_ERR_UNSUPPORTED,
                    "fused STFT+SCM: 1..8 channels per group (1..4 with two masks or n_fft = 1024)");
    if (!x || !Y) return fail(DISCO_ERR_INVALID, "null pointer");
    Tables tb;
    int rc = get_tables(n_fft, &tb);
    if (rc) return rc;
    const int T = disco_n_frames(length, n_fft);
    const int n_grp = (n_sig + C - 1) / C;
    StftArgs a;
    memset(&a, 0, sizeof(a));
    a.x = x;
    a.mask = mask;
    a.mask2 = mask2;
    a.Y = (float2*)Y;
    a.part = (float*)workspace;
    a.twiddle = tb.twiddle;
    a.window = tb.win_half;
    a.n_sig = n_sig;
    a.n_grp = n_grp;
    a.L = length;
    a.T = T;
    a.mask_ft = (mask_layout == DISCO_LAYOUT_FT);
    a.use_tma = (length % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    const StftPlan pl = plan_stft(n_grp, C, T, n_fft);
    a.slots_per_grp = pl.slots_per_grp;
    cudaStream_t st = (cudaStream_t)stream;
    if (n_mask > 0) {
        if (!mask || (n_mask == 2 && !mask2) || (!Rss) != (!Rnn)) return fail(DISCO_ERR_INVALID, "null pointer");
        const size_t need = stft_ws_bytes(n_grp, C, length, n_fft, n_mask);
        if (!workspace || workspace_bytes < need) return fail(DISCO_ERR_WORKSPACE, "workspace too small");
    }
    CU(launch_stft_scm(a, n_fft, C, pl.n_cta, n_mask, st), "stft_scm launch");
    if (n_mask == 1 && Rss)
        CU(launch_scm_finalize(a.part, (float2*)Rss, (float2*)Rnn, n_grp, pl.slots_per_grp, pl.tiles_per_grp,
                               pl.n_cta, C, n_fft / 2 + 1, T, 1, 0, st),
           "scm_finalize launch");
    return 0;
}

}  // namespace


extern "C" {

int disco_abi_version(void) { return DISCO_ABI_VERSION; }
const char* disco_last_error(void) { return g_err.c_str(); }

int disco_n_frames(int length, int n_fft) { return 1 + length / (n_fft / 2); }

int disco_set_reserved_sms(int n) {
    if (n < 0 || n > 64) return fail(DISCO_ERR_INVALID, "reserved SMs must be in 0..64");
    g_reserved_sms = n;
    return 0;
}

int disco_init(int n_fft) {
    if (!valid_nfft(n_fft)) return fail(DISCO_ERR_INVALID, "n_fft must be 256, 512 or 1024");
    Tables tb;
    return get_tables(n_fft, &tb);
}

int disco_stft(const float* x, void* Y, int n_sig, int length, int n_fft, void* stream) {
    // plain STFT: signals are grouped by 4 only to share a CTA's tile; groups are independent
    const int C = n_sig >= 4 ? 4 : n_sig;
    return stft_common(x, nullptr, nullptr, 0, Y, nullptr, nullptr, n_sig, C, length, n_fft, nullptr, 0, 0, stream);
}

size_t disco_stft_scm_workspace(int n_grp, int C, int length, int n_fft) {
    return stft_ws_bytes(n_grp, C, length, n_fft, 1);
}

int disco_stft_scm_supported(int n_fft, int C, int n_mask) { return stft_scm_supported(n_fft, C, n_mask) ? 1 : 0; }

int disco_stft_scm(const float* x, const float* mask, int mask_layout, void* Y, void* Rss, void* Rnn, int n_grp,
                   int C, int length, int n_fft, void* workspace, size_t workspace_bytes, void* stream) {
    if (n_grp <= 0) return fail(DISCO_ERR_INVALID, "n_grp must be positive");
    return stft_common(x, mask, nullptr, mask_layout, Y, Rss, Rnn, n_grp * C, C, length, n_fft, workspace,
                       workspace_bytes, 1, stream);
}

size_t disco_stft_scm2_workspace(int n_grp, int C, int length, int n_fft) {
    return stft_ws_bytes(n_grp, C, length, n_fft, 2);
}

int disco_stft_scm2(const float* x, const float* mask_a, const float* mask_b, int mask_layout, void* Y, int n_grp,
                    int C, int length, int n_fft, void* workspace, size_t workspace_bytes, void* stream) {
    if (n_grp <= 0) return fail(DISCO_ERR_INVALID, "n_grp must be positive");
    return stft_common(x, mask_a, mask_b, mask_layout, Y, nullptr, nullptr, n_grp * C, C, length, n_fft, workspace,
                       workspace_bytes, 2, stream);
}

int disco_scm_from_workspace(const void* workspace, int n_set, int set, void* Rss, void* Rnn, int n_grp, int C,
                             int length, int n_fft, void* stream) {
    if (!valid_nfft(n_fft) || n_grp < 1 || C < 1 || C > 8 || n_set < 1 || n_set > 2 || set < 0 || set >= n_set ||
        !workspace || !Rss || !Rnn)
        return fail(DISCO_ERR_INVALID, "bad arguments");
    const int T = disco_n_frames(length, n_fft);
    const StftPlan pl = plan_stft(n_grp, C, T, n_fft);
    CU(launch_scm_finalize((const float*)workspace, (float2*)Rss, (float2*)Rnn, n_grp, pl.slots_per_grp,
                           pl.tiles_per_grp, pl.n_cta, C, n_fft / 2 + 1, T, n_set, set, (cudaStream_t)stream),
       "scm_finalize launch");
    return 0;
}

int disco_tf_mask(const void* S, const void* N, float* M, size_t n_elem, int kind, int power, float bin_thr_db,
                  void* stream) {
    if (kind < 0 || kind > 2) return fail(DISCO_ERR_INVALID, "unknown mask kind");
    if (power < 0 || power > 9) return fail(DISCO_ERR_INVALID, "mask power must be a single digit");
    if (!S || !N || !M) return fail(DISCO_ERR_INVALID, "null pointer");
    const float thr = powf(10.f, bin_thr_db / 10.f);  // math_utils.db2lin
    CU(launch_tf_mask((const float2*)S, (const float2*)N, M, n_elem, kind, power, thr, (cudaStream_t)stream),
       "tf_mask launch");
    return 0;
}

static int make_cat(CatArgs* c, const void* Y, const void* Z, int n_utt, int K, int C, int T, int n_fft,
                    const int* node_sel, int n_sel, int z_layout = DISCO_Z_UTT_MAJOR) {
    if (!valid_nfft(n_fft)) return fail(DISCO_ERR_INVALID, "n_fft must be 256, 512 or 1024");
    if (z_layout != DISCO_Z_UTT_MAJOR && z_layout != DISCO_Z_NODE_MAJOR) return fail(DISCO_ERR_INVALID, "bad z_layout");
    if (n_utt <= 0 || K < 1 || C < 1 || T < 1) return fail(DISCO_ERR_INVALID, "bad sizes");
    if (C + K - 1 > 16) return fail(DISCO_ERR_UNSUPPORTED, "C + K - 1 must be <= 16");
    if (!Y || (K > 1 && !Z)) return fail(DISCO_ERR_INVALID, "null pointer");
    c->Y = (const float2*)Y;
    c->Z = (const float2*)Z;
    c->z_sb = (z_layout == DISCO_Z_NODE_MAJOR) ? 1 : K;
    c->z_sk = (z_layout == DISCO_Z_NODE_MAJOR) ? n_utt : 1;
    c->C = C;
    c->K = K;
    c->T = T;
    c->F = n_fft / 2 + 1;
    if (K > 16) return fail(DISCO_ERR_UNSUPPORTED, "at most 16 nodes");
    if (node_sel) {
        if (n_sel < 1 || n_sel > K) return fail(DISCO_ERR_INVALID, "bad node selection");
        for (int i = 0; i < n_sel; ++i) {
            if (node_sel[i] < 0 || node_sel[i] >= K || (i > 0 && node_sel[i] <= node_sel[i - 1]))
                return fail(DISCO_ERR_INVALID, "node selection must be ascending node indices");
            c->sel[i] = node_sel[i];
        }
        c->n_sel = n_sel;
    } else {
        c->n_sel = K;
        for (int i = 0; i < K; ++i) c->sel[i] = i;
    }
    c->n_grp = n_utt * c->n_sel;
    return 0;
}

int disco_masked_scm(const void* Y, const void* Z, const float* mask, int mask_layout, void* Rss, void* Rnn,
                     int n_utt, int K, int C, int T, int n_fft, const int* node_sel, int n_sel, int z_layout,
                     void* stream) {
    ScmArgs a;
    memset(&a, 0, sizeof(a));
    int rc = make_cat(&a.in, Y, Z, n_utt, K, C, T, n_fft, node_sel, n_sel, z_layout);
    if (rc) return rc;
    if (!Rss || !Rnn) return fail(DISCO_ERR_INVALID, "null pointer");
    if (a.in.n_grp > kMaxGridYZ) return fail(DISCO_ERR_UNSUPPORTED, "at most 65535 (utterance, node) groups per call");
    a.mask = mask;
    a.mask_ft = (mask_layout == DISCO_LAYOUT_FT);
    a.Rss = (float2*)Rss;
    a.Rnn = (float2*)Rnn;
    CU(launch_masked_scm(a, (cudaStream_t)stream), "masked_scm launch");
    return 0;
}

int disco_filter_sum_scm(const void* W1, const void* Y, const float* mask, int mask_layout, void* z_out, void* zn_out,
                         int ref, void* Rss, void* Rnn, int n_grp, int C, int T, int n_fft, void* stream) {
    ScmArgs a;
    memset(&a, 0, sizeof(a));
    int rc = make_cat(&a.in, Y, nullptr, n_grp, 1, C, T, n_fft, nullptr, 0);
    if (rc) return rc;
    if (!W1 || !z_out || !Rss || !Rnn || !mask) return fail(DISCO_ERR_INVALID, "null pointer");
    if (ref < 0 || ref >= C) return fail(DISCO_ERR_INVALID, "ref channel out of range");
    a.mask = mask;
    a.mask_ft = (mask_layout == DISCO_LAYOUT_FT);
    a.Rss = (float2*)Rss;
    a.Rnn = (float2*)Rnn;
    a.W1 = (const float2*)W1;
    a.z_out = (float2*)z_out;
    a.zn_out = (float2*)zn_out;
    a.ref = ref;
    CU(launch_masked_scm(a, (cudaStream_t)stream), "filter_sum_scm launch");
    return 0;
}

int disco_tango_mid_supported(int C, int K) { return tango_mid_supported(C, K) ? 1 : 0; }

int disco_tango_mid(const void* W1, const void* Y, const float* mask_w, void* Z, void* ZN, int ref, void* Rss,
                    void* Rnn, int n_utt, int K, int C, int T, int n_fft, void* stream) {
    if (!valid_nfft(n_fft)) return fail(DISCO_ERR_INVALID, "n_fft must be 256, 512 or 1024");
    if (!W1 || !Y || !mask_w || !Z || !Rss || !Rnn || n_utt < 1 || T < 1)
        return fail(DISCO_ERR_INVALID, "bad arguments");
    if (!tango_mid_supported(C, K)) return fail(DISCO_ERR_UNSUPPORTED, "no fused middle pass for this (C, K)");
    if (ref < 0 || ref >= C) return fail(DISCO_ERR_INVALID, "ref channel out of range");
    MidArgs a;
    a.Y = (const float2*)Y;
    a.W1 = (const float2*)W1;
    a.mask = mask_w;
    a.Z = (float2*)Z;
    a.ZN = (float2*)ZN;
    a.Rss = (float2*)Rss;
    a.Rnn = (float2*)Rnn;
    a.B = n_utt;
    a.K = K;
    a.C = C;
    a.T = T;
    a.F = n_fft / 2 + 1;
    a.ref = ref;
    CU(launch_tango_mid(a, (cudaStream_t)stream), "tango_mid launch");
    return 0;
}

static int solve_workspace(const void* workspace, int n_set, void* W, void* T1, void* Rss, void* Rnn, int n_grp, int C,
                           int length, int n_fft, int filter_type, int rank, double mu, void* stream) {
    if (filter_type < 0 || filter_type > 2) return fail(DISCO_ERR_INVALID, "Unknown filter reference");
    if (!valid_nfft(n_fft) || n_grp < 1 || C < 1 || C > 4 || n_set < 1 || n_set > 2 || !workspace || !W ||
        (!Rss) != (!Rnn))
        return fail(DISCO_ERR_INVALID, "bad arguments");
    const int T = disco_n_frames(length, n_fft), F = n_fft / 2 + 1;
    const StftPlan pl = plan_stft(n_grp, C, T, n_fft);
    SolveArgs a;
    memset(&a, 0, sizeof(a));
    a.Rss = (const float2*)Rss;
    a.Rnn = (const float2*)Rnn;
    a.W = (float2*)W;
    a.T1 = (float2*)T1;
    a.n_mat = n_set * n_grp * F;
    a.n_set = n_set;
    a.D = C;
    a.type = filter_type;
    a.rank = rank;
    a.mu = mu;
    a.part = (const float*)workspace;
    a.slots_per_grp = pl.slots_per_grp;
    a.tiles_per_grp = pl.tiles_per_grp;
    a.n_cta = pl.n_cta;
    a.F = F;
    a.inv_T = 1.0f / (float)T;
    CU(launch_mwf_solve(a, (cudaStream_t)stream), "mwf_solve launch");
    return 0;
}

int disco_mwf_solve_workspace(const void* workspace, void* W, void* T1, void* Rss, void* Rnn, int n_grp, int C,
                              int length, int n_fft, int filter_type, int rank, double mu, void* stream) {
    return solve_workspace(workspace, 1, W, T1, Rss, Rnn, n_grp, C, length, n_fft, filter_type, rank, mu, stream);
}

int disco_mwf_solve_workspace2(const void* workspace, void* W, void* T1, int n_grp, int C, int length, int n_fft,
                               int filter_type, int rank, double mu, void* stream) {
    return solve_workspace(workspace, 2, W, T1, nullptr, nullptr, n_grp, C, length, n_fft, filter_type, rank, mu,
                           stream);
}

int disco_mwf_solve(const void* Rss, const void* Rnn, void* W, void* T1, int n_mat, int D, int filter_type,
                    int rank, double mu, void* stream) {
    if (filter_type < 0 || filter_type > 2) return fail(DISCO_ERR_INVALID, "Unknown filter reference");
    if (D < 1 || D > 16) return fail(DISCO_ERR_UNSUPPORTED, "D must be in 1..16");
    if (n_mat < 0 || !Rss || !Rnn || !W) return fail(DISCO_ERR_INVALID, "bad arguments");
    SolveArgs a;
    memset(&a, 0, sizeof(a));
    a.Rss = (const float2*)Rss;
    a.Rnn = (const float2*)Rnn;
    a.W = (float2*)W;
    a.T1 = (float2*)T1;
    a.n_mat = n_mat;
    a.D = D;
    a.type = filter_type;
    a.rank = rank;
    a.mu = mu;
    CU(launch_mwf_solve(a, (cudaStream_t)stream), "mwf_solve launch");
    return 0;
}

int disco_filter_sum(const void* W, int conj_w, const void* Y, const void* Z, void* out, void* resid, int ref,
                     int out_layout, int n_utt, int K, int C, int T, int n_fft, const int* node_sel, int n_sel,
                     int z_layout, void* stream) {
    FilterArgs a;
    memset(&a, 0, sizeof(a));
    int rc = make_cat(&a.in, Y, Z, n_utt, K, C, T, n_fft, node_sel, n_sel, z_layout);
    if (rc) return rc;
    if (!W || !out) return fail(DISCO_ERR_INVALID, "null pointer");
    if (ref < 0 || ref >= C + K - 1) return fail(DISCO_ERR_INVALID, "ref channel out of range");
    if (a.in.n_grp > kMaxGridYZ) return fail(DISCO_ERR_UNSUPPORTED, "at most 65535 (utterance, node) groups per call");
    a.W = (const float2*)W;
    a.conj_w = conj_w;
    a.out = (float2*)out;
    a.resid = (float2*)resid;
    a.ref = ref;
    a.out_ft = (out_layout == DISCO_LAYOUT_FT);
    CU(launch_filter_sum(a, (cudaStream_t)stream), "filter_sum launch");
    return 0;
}

int disco_filter_dual(const void* W1, const void* W2, const void* Y, void* z, void* zn, void* yf, int ref,
                      int out_layout, int n_grp, int C, int T, int n_fft, void* stream) {
    if (!valid_nfft(n_fft)) return fail(DISCO_ERR_INVALID, "n_fft must be 256, 512 or 1024");
    if (!W1 || !W2 || !Y || !z || !yf || n_grp < 1 || T < 1) return fail(DISCO_ERR_INVALID, "bad arguments");
    if (C < 1 || C > 4) return fail(DISCO_ERR_UNSUPPORTED, "filter_dual supports 1..4 channels");
    if (ref < 0 || ref >= C) return fail(DISCO_ERR_INVALID, "ref channel out of range");
    if (n_grp > 65535) return fail(DISCO_ERR_UNSUPPORTED, "at most 65535 groups per call");
    DualFilterArgs a;
    a.Y = (const float2*)Y;
    a.W1 = (const float2*)W1;
    a.W2 = (const float2*)W2;
    a.z = (float2*)z;
    a.zn = (float2*)zn;
    a.yf = (float2*)yf;
    a.n_grp = n_grp;
    a.C = C;
    a.T = T;
    a.F = n_fft / 2 + 1;
    a.ref = ref;
    a.out_ft = (out_layout == DISCO_LAYOUT_FT);
    CU(launch_filter_dual(a, sm_count(), (cudaStream_t)stream), "filter_dual launch");
    return 0;
}

int disco_istft(const void* Y, float* x, int n_sig, int T, int length, int n_fft, void* stream) {
    if (!valid_nfft(n_fft)) return fail(DISCO_ERR_INVALID, "n_fft must be 256, 512 or 1024");
    if (n_sig <= 0 || T < 1 || length < 1 || !Y || !x) return fail(DISCO_ERR_INVALID, "bad arguments");
    Tables tb;
    int rc = get_tables(n_fft, &tb);
    if (rc) return rc;
    IstftArgs a;
    a.Y = (const float2*)Y;
    a.x = x;
    a.twiddle = tb.twiddle;
    a.window = tb.win;
    a.n_sig = n_sig;
    a.L = length;
    a.T = T;
    CU(launch_istft(a, n_fft, (cudaStream_t)stream), "istft launch");
    return 0;
}

int disco_scm_recursive(const void* Y, const void* Z, const float* mask, const void* R0ss, const void* R0nn, void* Rss,
                        void* Rnn, double lambda_cor, int block, int weight_power, int n_utt, int K, int C, int T,
                        int n_fft, const int* node_sel, int n_sel, void* stream) {
    OnlineArgs a;
    memset(&a, 0, sizeof(a));
    int rc = make_cat(&a.in, Y, Z, n_utt, K, C, T, n_fft, node_sel, n_sel);
    if (rc) return rc;
    if (C + K - 1 > 8) return fail(DISCO_ERR_UNSUPPORTED, "recursive SCM: C + K - 1 must be <= 8");
    if (a.in.n_grp > kMaxGridYZ) return fail(DISCO_ERR_UNSUPPORTED, "at most 65535 (utterance, node) groups per call");
    if (!Rss || !Rnn || (!R0ss) != (!R0nn)) return fail(DISCO_ERR_INVALID, "null pointer");
    if (block < 1 || block > 64) return fail(DISCO_ERR_INVALID, "block must be 1..64 frames");
    if (!(lambda_cor >= 0.0 && lambda_cor < 1.0)) return fail(DISCO_ERR_INVALID, "lambda_cor must be in [0, 1)");
    if (weight_power != 1 && weight_power != 2) return fail(DISCO_ERR_INVALID, "weight_power must be 1 or 2");
    a.mask = mask;
    a.R0ss = (const float2*)R0ss;
    a.R0nn = (const float2*)R0nn;
    a.Rss = (float2*)Rss;
    a.Rnn = (float2*)Rnn;
    a.P = block;
    a.J = (T + block - 1) / block;
    a.power = weight_power;
    double g = 1.0 - lambda_cor;
    for (int k = 0; k < 64; ++k) {
        a.gw[k] = (float)g;
        g *= lambda_cor;
    }
    a.lam_block = (float)pow(lambda_cor, block);
    a.lam_last = (float)pow(lambda_cor, T - (a.J - 1) * block);
    CU(launch_scm_recursive(a, (cudaStream_t)stream), "scm_recursive launch");
    return 0;
}

int disco_filter_sum_blocks(const void* W, int conj_w, const void* Y, const void* Z, void* out, void* resid, int ref,
                            int block, int lag, int n_utt, int K, int C, int T, int n_fft, const int* node_sel,
                            int n_sel, void* stream) {
    OnlineFilterArgs a;
    memset(&a, 0, sizeof(a));
    int rc = make_cat(&a.in, Y, Z, n_utt, K, C, T, n_fft, node_sel, n_sel);
    if (rc) return rc;
    if (C + K - 1 > 8) return fail(DISCO_ERR_UNSUPPORTED, "block filter: C + K - 1 must be <= 8");
    if (a.in.n_grp > kMaxGridYZ) return fail(DISCO_ERR_UNSUPPORTED, "at most 65535 (utterance, node) groups per call");
    if (!W || !out) return fail(DISCO_ERR_INVALID, "null pointer");
    if (block < 1 || block > 64 || lag < 0) return fail(DISCO_ERR_INVALID, "bad block / lag");
    if (ref < 0 || ref >= C + K - 1) return fail(DISCO_ERR_INVALID, "ref channel out of range");
    a.W = (const float2*)W;
    a.conj_w = conj_w;
    a.out = (float2*)out;
    a.resid = (float2*)resid;
    a.ref = ref;
    a.P = block;
    a.J = (T + block - 1) / block;
    a.lag = lag;
    CU(launch_filter_sum_blocks(a, (cudaStream_t)stream), "filter_sum_blocks launch");
    return 0;
}

int disco_band_stats(const float* x, const float* sel, const double* ba, double* stats, int n_sig, int length,
                     long long row_stride, int n_band, int order, void* stream) {
    if (!x || !ba || !stats || n_sig < 1 || length < 1 || n_band < 1 || row_stride < length)
        return fail(DISCO_ERR_INVALID, "bad arguments");
    if (order != 2 && order != 4 && order != 8 && order != 16)
        return fail(DISCO_ERR_UNSUPPORTED, "filter order must be 2, 4, 8 or 16");
    BankArgs a;
    a.x = x;
    a.sel = sel;
    a.ba = ba;
    a.stats = stats;
    a.n_sig = n_sig;
    a.L = length;
    a.n_band = n_band;
    a.ldx = row_stride;
    CU(launch_band_stats(a, order, (cudaStream_t)stream), "band_stats launch");
    return 0;
}

int disco_transpose_c64(const void* in, void* out, int batch, int rows, int cols, void* stream) {
    if (!in || !out) return fail(DISCO_ERR_INVALID, "null pointer");
    if (batch > kMaxGridYZ) return fail(DISCO_ERR_UNSUPPORTED, "at most 65535 planes per call");
    CU(launch_transpose_c64((const float2*)in, (float2*)out, batch, rows, cols, (cudaStream_t)stream),
       "transpose launch");
    return 0;
}
int disco_transpose_f32(const float* in, float* out, int batch, int rows, int cols, void* stream) {
    if (!in || !out) return fail(DISCO_ERR_INVALID, "null pointer");
    if (batch > kMaxGridYZ) return fail(DISCO_ERR_UNSUPPORTED, "at most 65535 planes per call");
    CU(launch_transpose_f32(in, out, batch, rows, cols, (cudaStream_t)stream), "transpose launch");
    return 0;
}
int disco_apply_mask(const void* in, const float* m, void* out, size_t n_elem, int one_minus, void* stream) {
    if (!in || !m || !out) return fail(DISCO_ERR_INVALID, "null pointer");
    CU(launch_apply_mask((const float2*)in, m, (float2*)out, n_elem, n_elem, 1, one_minus, (cudaStream_t)stream),
       "apply_mask launch");
    return 0;
}
int disco_apply_mask_channels(const void* in, const float* m, void* out, size_t n_grp, int chans, size_t plane,
                              int one_minus, void* stream) {
    if (!in || !m || !out || chans < 1) return fail(DISCO_ERR_INVALID, "bad arguments");
    CU(launch_apply_mask((const float2*)in, m, (float2*)out, n_grp * chans * plane, plane, chans, one_minus,
                         (cudaStream_t)stream),
       "apply_mask launch");
    return 0;
}

}  // extern "C"
