// extern "C" surface of libgigagan_sm100.so (declared in include/gigagan_sm100.h).
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include "../../include/gigagan_sm100.h"
#include "gg_internal.h"

static thread_local char g_err[512] = "";

int gg_fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return -1;
}
int gg_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return 0;
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return -2;
}

#define ST ((cudaStream_t)stream)
static int g_flags = 0;
static int dbg_fallback() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GG_DEBUG_FALLBACK"); v = (e && e[0] == '1') ? 1 : 0; }
  return v;
}

extern "C" {
const char* gg_last_error(void) { return g_err; }
int gg_version(void) { return 100; }
int gg_set_flags(int flags) { int o = g_flags; g_flags = flags; return o; }
int gg_has_tcgen05(void) {
#ifdef GG_NO_TC
  return 0;
#else
  return 1;
#endif
}

static int conv_fprop_any(const void* x, const void* w, const float* bias, const void* res, void* y, int N, int H, int W,
                          int Cin, int OH, int OW, int Cout, int KH, int KW, int stride, int pad, int per_sample_w, int act,
                          float gain, const long* ystr, int dtype, cudaStream_t st) {
#ifndef GG_NO_TC
  if (dtype == GG_BF16 && !(g_flags & 1)) {
    int r = 1;
    if (!ystr && !(g_flags & 2))           // thin 128^2 / 256^2 layers: input rows staged once, all taps from shared memory
      r = ggi_tc_conv_thin(x, w, bias, res, y, N, H, W, Cin, OH, OW, Cout, KH, KW, stride, pad, per_sample_w, act, gain, st);
    if (r <= 0) return r;
    r = ggi_tc_conv_fprop(x, w, bias, res, y, N, H, W, Cin, OH, OW, Cout, KH, KW, stride, pad, per_sample_w, act, gain, ystr, st);
    if (r <= 0) return r;
    if (dbg_fallback()) fprintf(stderr, "[gg] FFMA fprop: N%d H%d W%d Cin%d -> OH%d OW%d Cout%d k%dx%d s%d p%d ps%d\n", N, H, W, Cin, OH, OW, Cout, KH, KW, stride, pad, per_sample_w);
  }
#endif
  return ggi_simt_conv_fprop(x, w, bias, res, y, N, H, W, Cin, OH, OW, Cout, KH, KW, stride, pad, per_sample_w, act, gain, ystr, dtype, st);
}
int gg_conv2d_fprop(const void* x, const void* w, const float* bias, const void* res, void* y, int N, int H, int W,
                    int Cin, int OH, int OW, int Cout, int KH, int KW, int stride, int pad, int per_sample_w, int act,
                    float gain, int dtype, gg_stream_t stream) {
  return conv_fprop_any(x, w, bias, res, y, N, H, W, Cin, OH, OW, Cout, KH, KW, stride, pad, per_sample_w, act, gain, nullptr, dtype, ST);
}
int gg_conv2d_fprop_strided(const void* x, const void* w, const float* bias, void* y, int N, int H, int W, int Cin,
                            int OH, int OW, int Cout, int KH, int KW, int stride, int pad, int per_sample_w, int act,
                            float gain, int64_t y_off, int64_t y_sn, int64_t y_sh, int64_t y_sw, int dtype,
                            gg_stream_t stream) {
  long ystr[4] = {(long)y_off, (long)y_sn, (long)y_sh, (long)y_sw};
  return conv_fprop_any(x, w, bias, nullptr, y, N, H, W, Cin, OH, OW, Cout, KH, KW, stride, pad, per_sample_w, act, gain, ystr, dtype, ST);
}
int gg_conv2d_dgrad(const void* dy, const void* w, void* dx, int N, int H, int W, int Cin, int OH, int OW, int Cout,
                    int KH, int KW, int stride, int pad, int per_sample_w, int dtype, gg_stream_t stream) {
  return ggi_simt_conv_dgrad(dy, w, dx, N, H, W, Cin, OH, OW, Cout, KH, KW, stride, pad, per_sample_w, dtype, ST);
}
int gg_conv2d_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int OH, int OW, int Cout,
                    int KH, int KW, int stride, int pad, int per_sample_w, int dtype, gg_stream_t stream) {
#ifndef GG_NO_TC
  if (dtype == GG_BF16 && !(g_flags & 1)) {
    int r = 1;
    if (!(g_flags & 6)) r = ggi_tc_conv_thin_wgrad(x, dy, dw, N, H, W, Cin, OH, OW, Cout, KH, KW, stride, pad, per_sample_w, ST);
    if (r <= 0) return r;
    r = ggi_tc_conv_wgrad(x, dy, dw, N, H, W, Cin, OH, OW, Cout, KH, KW, stride, pad, per_sample_w, ST);
    if (r <= 0) return r;
    if (dbg_fallback()) fprintf(stderr, "[gg] FFMA wgrad: N%d H%d W%d Cin%d -> OH%d OW%d Cout%d k%dx%d s%d p%d ps%d\n", N, H, W, Cin, OH, OW, Cout, KH, KW, stride, pad, per_sample_w);
  }
#endif
  return ggi_simt_conv_wgrad(x, dy, dw, N, H, W, Cin, OH, OW, Cout, KH, KW, stride, pad, per_sample_w, dtype, ST);
}
int gg_bmm(const void* A, const void* B, const float* bias, void* C, int b1, int b2, int M, int N, int K,
           const int64_t* h_sa, const int64_t* h_sb, const int64_t* h_sc, float alpha, int dtype, gg_stream_t stream) {
#ifndef GG_NO_TC
  if (dtype == GG_BF16 && !(g_flags & 1)) {
    int r = ggi_tc_bmm(A, B, bias, C, b1, b2, M, N, K, (const long*)h_sa, (const long*)h_sb, (const long*)h_sc, alpha, ST);
    if (r <= 0) return r;
    if (dbg_fallback()) fprintf(stderr, "[gg] FFMA bmm: b%dx%d M%d N%d K%d sa(%ld,%ld,%ld,%ld) sb(%ld,%ld,%ld,%ld) A%%16=%d B%%16=%d\n", b1, b2, M, N, K,
                                (long)h_sa[0], (long)h_sa[1], (long)h_sa[2], (long)h_sa[3], (long)h_sb[0], (long)h_sb[1], (long)h_sb[2], (long)h_sb[3],
                                (int)((uintptr_t)A & 15), (int)((uintptr_t)B & 15));
  }
#endif
  return ggi_simt_bmm(A, B, bias, C, b1, b2, M, N, K, (const long*)h_sa, (const long*)h_sb, (const long*)h_sc, alpha, dtype, ST);
}
int gg_pw_unary(int kind, int level, const void* x, const void* a, const void* b, void* out, int64_t n, int dtype, gg_stream_t stream) {
  return ggi_pw_unary(kind, level, x, a, b, out, n, dtype, ST);
}
int gg_pw_mul(const void* a, const void* b, void* out, int64_t n, int dtype, gg_stream_t stream) { return ggi_pw_mul(a, b, out, n, dtype, ST); }
int gg_pw_axpby(float alpha, const void* x, float beta, const void* y, void* out, int64_t n, int dtype, gg_stream_t stream) {
  return ggi_pw_axpby(alpha, x, beta, y, out, n, dtype, ST);
}
int gg_pw_bcast(const void* x, const float* s, void* out, int64_t R, int C, int P, int Ns, int mode, int op, int dtype, gg_stream_t stream) {
  return ggi_pw_bcast(x, s, out, R, C, P, Ns, mode, op, dtype, ST);
}
int gg_red_rowdot(const void* a, const void* b, float* out, int64_t R, int C, int dtype, gg_stream_t stream) { return ggi_red_rowdot(a, b, out, R, C, dtype, ST); }
int gg_red_dot_sc(const void* a, const void* b, float* out, int64_t R, int C, int P, int Ns, int dtype, gg_stream_t stream) {
  return ggi_red_dot_sc(a, b, out, R, C, P, Ns, dtype, ST);
}
int gg_softmax_rows(const void* s, const float* bias, void* p, int64_t R, int C, int P, int Ns, int dtype, gg_stream_t stream) {
  return ggi_softmax_rows(s, bias, p, R, C, P, Ns, dtype, ST);
}
int gg_softmax_bwd_rows(const void* p, const void* gp, void* ds, int64_t R, int C, int dtype, gg_stream_t stream) {
  return ggi_softmax_bwd_rows(p, gp, nullptr, ds, R, C, dtype, ST);
}
int gg_softmax_bwd_rows_add(const void* p, const void* gp, const void* gp2, void* ds, int64_t R, int C, int dtype,
                            gg_stream_t stream) {
  return ggi_softmax_bwd_rows(p, gp, gp2, ds, R, C, dtype, ST);
}
int gg_softmax_bwd2_rows(const void* p, const void* gp, const void* G, void* d_p, void* d_gp, int64_t R, int C, int dtype,
                         gg_stream_t stream) {
  int r = ggi_softmax_bwd2_rows(p, gp, G, d_p, d_gp, R, C, dtype, ST);
  if (r == 1) return gg_fail("gg_softmax_bwd2_rows: row length %d not supported by the one-pass kernel", C);
  return r;
}
int gg_resample2d(const void* x, void* y, int N, int H, int W, int C, int OH, int OW, const int* iy, const float* wy,
                  int Ty, const int* ix, const float* wx, int Tx, int dtype, gg_stream_t stream) {
  return ggi_resample2d(x, y, N, H, W, C, OH, OW, iy, wy, Ty, ix, wx, Tx, dtype, ST);
}
int gg_nchw_to_nhwc(const float* src, void* dst, int N, int C, int HW, int Cpad, int dtype, gg_stream_t stream) { return ggi_nchw_to_nhwc(src, dst, N, C, HW, Cpad, dtype, ST); }
int gg_nhwc_to_nchw(const void* src, float* dst, int N, int C, int HW, int Cpad, int dtype, gg_stream_t stream) { return ggi_nhwc_to_nchw(src, dst, N, C, HW, Cpad, dtype, ST); }
int gg_noise_act_fwd(const void* x, const float* noise, const float* wn, void* y, int64_t R, int C, int dtype, gg_stream_t stream) {
  return ggi_noise_act_fwd(x, noise, wn, y, R, C, dtype, ST);
}
int gg_noise_act_bwd(const void* y, const void* gy, const float* noise, void* dx, float* dwn, int64_t R, int C, int dtype, gg_stream_t stream) {
  return ggi_noise_act_bwd(y, gy, noise, dx, dwn, R, C, dtype, ST);
}
int gg_adaconv_weights_fwd(const float* bank, const float* mod, const float* kmod, void* w, float* attn, float* dinv,
                           int B, int n, int O, int I, int KK, int demod, float eps, int Opad, int64_t mod_ld,
                           int64_t kmod_ld, int dtype, gg_stream_t stream) {
  return ggi_adaconv_weights_fwd(bank, mod, kmod, w, attn, dinv, B, n, O, I, KK, demod, eps, Opad, (long)mod_ld,
                                 (long)kmod_ld, dtype, ST);
}
int gg_adaconv_weights_bwd(const float* bank, const float* mod, const float* attn, const float* dinv, const float* gw,
                           float* dbank, float* dmod, float* dkmod, float* gattn_ws, int B, int n, int O, int I, int KK,
                           int demod, float eps, int Opad, int64_t mod_ld, gg_stream_t stream) {
  return ggi_adaconv_weights_bwd(bank, mod, attn, dinv, gw, dbank, dmod, dkmod, gattn_ws, B, n, O, I, KK, demod, eps, Opad,
                                 (long)mod_ld, nullptr, nullptr, ST);
}
int gg_sbank_bwd_stats(const float* bank, const float* mod, const float* attn, const float* dinv, const float* gdinv,
                       const float* dw_add, float* dbank, float* dmod, float* dkmod, float* gattn_ws, int B, int n, int O,
                       int I, int KK, float eps, int64_t mod_ld, gg_stream_t stream) {
  return ggi_adaconv_weights_bwd(bank, mod, attn, dinv, nullptr, dbank, dmod, dkmod, gattn_ws, B, n, O, I, KK, 1, eps, O,
                                 (long)mod_ld, gdinv, dw_add, ST);
}
int gg_red_dot_sc_acc(const void* a, const void* b, float* out, int64_t R, int C, int P, int Ns, int dtype, gg_stream_t stream) {
  return ggi_red_dot_sc_acc(a, b, out, R, C, P, Ns, 1, dtype, ST);
}
int gg_attn_fwd(const void* q, const void* k, const void* v, const float* null_kv, void* o, float* lse, float* ksq_ws,
                int B, int heads, int nq, int nk, int d, int64_t q_rs, int64_t k_rs, int64_t v_rs, int64_t o_rs, float scale,
                int mode, int dtype, gg_stream_t stream) {
#ifndef GG_NO_TC
  if (dtype == GG_BF16 && !(g_flags & 1)) {
    // flag 8: first-generation kernels (8 softmax warps, chunked TMEM reads).  Second generation: the forward runs two
    // CTAs per SM with 8 softmax warps each, the backward 16 warps; flags 16 / 32: one CTA per SM with 8 / 16 warps in every
    // kernel, flag 128: one-CTA forward with 8 warps, flag 64: two-pass L2 forward (A/B measurements).
    int r = (g_flags & 8) ? ggi_tc_attn_fwd(q, k, v, null_kv, o, lse, ksq_ws, B, heads, nq, nk, d, q_rs, k_rs, v_rs, o_rs, scale, mode, ST)
                          : ggi_tc2_attn_fwd(q, k, v, null_kv, o, lse, ksq_ws, B, heads, nq, nk, d, q_rs, k_rs, v_rs, o_rs, scale, mode,
                                             ((g_flags & 32) ? 16 : 8) | (g_flags & 192) | ((g_flags & 48) ? 256 : 0), ST);
    if (r <= 0) return r;
  }
#endif
  return ggi_attn_fwd(q, k, v, null_kv, o, lse, B, heads, nq, nk, d, q_rs, k_rs, v_rs, o_rs, scale, mode, dtype, ST);
}
int gg_attn_bwd(const void* q, const void* k, const void* v, const float* null_kv, const void* o, const void* go,
                const float* lse, void* dq, void* dk, void* dv, float* dnull_kv, float* delta_ws, float* ksq_ws, int B,
                int heads, int nq, int nk, int d, int64_t q_rs, int64_t k_rs, int64_t v_rs, int64_t o_rs, float scale,
                int mode, int dtype, gg_stream_t stream) {
#ifndef GG_NO_TC
  if (dtype == GG_BF16 && !(g_flags & 1)) {
    int r = (g_flags & 8) ? ggi_tc_attn_bwd(q, k, v, null_kv, o, go, lse, dq, dk, dv, dnull_kv, delta_ws, ksq_ws, B, heads, nq, nk, d, q_rs, k_rs, v_rs, o_rs, scale, mode, ST)
                          : ggi_tc2_attn_bwd(q, k, v, null_kv, o, go, lse, dq, dk, dv, dnull_kv, delta_ws, ksq_ws, B, heads, nq, nk, d, q_rs, k_rs, v_rs, o_rs, scale, mode,
                                             (g_flags & 16) ? 8 : 16, ST);
    if (r <= 0) return r;
  }
#endif
  return ggi_attn_bwd(q, k, v, null_kv, o, go, lse, dq, dk, dv, dnull_kv, delta_ws, B, heads, nq, nk, d, q_rs, k_rs, v_rs, o_rs, scale, mode, dtype, ST);
}
int gg_adamw(float* p, const float* g, float* m, float* v, const void* chunks, int nchunks, const int* step_ptr, float lr,
             float b1, float b2, float eps, float wd, float grad_scale, gg_stream_t stream) {
  return ggi_adamw(p, g, m, v, chunks, nchunks, step_ptr, lr, b1, b2, eps, wd, grad_scale, ST);
}
int gg_incr(int* p, gg_stream_t stream) { return ggi_incr(p, ST); }
int gg_maxpool2_fwd(const void* x, void* y, int N, int H, int W, int C, int dtype, gg_stream_t stream) { return ggi_maxpool2_fwd(x, y, N, H, W, C, dtype, ST); }
int gg_maxpool2_bwd(const void* x, const void* gy, void* gx, int N, int H, int W, int C, int dtype, gg_stream_t stream) {
  return ggi_maxpool2_bwd(x, gy, gx, N, H, W, C, dtype, ST);
}
int gg_softmax_tokens(const void* x, void* y, int B, int n, int C, int dtype, gg_stream_t stream) { return ggi_softmax_tokens(x, y, B, n, C, dtype, ST); }
int gg_rmsnorm_fwd(const void* x, const float* gamma, void* y, float* inv, int64_t R, int C, float s, int dtype, gg_stream_t stream) {
  return ggi_rmsnorm_fwd(x, gamma, y, inv, (long)R, C, s, dtype, ST);
}
int gg_rmsnorm_bwd(const void* x, const float* gamma, const float* inv, const void* gy, void* gx, float* dgamma, int64_t R, int C,
                   float s, int dtype, gg_stream_t stream) {
  return ggi_rmsnorm_bwd(x, gamma, inv, gy, gx, dgamma, (long)R, C, s, dtype, ST);
}
int gg_debug_mma_chain(int N, int nacc, int iters, void* out, gg_stream_t stream) { return ggi_debug_mma_chain(N, nacc, iters, (unsigned long long*)out, ST); }
int gg_debug_thin_trace(void* buf) { return ggi_debug_thin_trace((unsigned long long*)buf); }
int gg_lrelu_bwd_bias(const void* y, const void* gy, void* out, float* dbias, int64_t R, int C, int accumulate, int dtype,
                      gg_stream_t stream) {
  return ggi_lrelu_bwd_bias(y, gy, out, dbias, (long)R, C, accumulate, dtype, ST);
}
int gg_wgrad_sink(const float* dw, float* dst, int O, int I, int KK, int Ipad, gg_stream_t stream) {
  return ggi_wgrad_sink(dw, dst, O, I, KK, Ipad, ST);
}
int gg_weight_prep_multi(const float* master, const void* entries, const void* chunks, int nchunks, void* fwd, void* bwd,
                         int dtype, gg_stream_t stream) {
  return ggi_weight_prep_multi(master, entries, chunks, nchunks, fwd, bwd, dtype, ST);
}
int gg_gan_loss_fwd(const void* const* h_x, const int64_t* h_meta, int k, int mode, float w_ms, float* out, gg_stream_t stream) {
  return ggi_gan_loss_fwd(h_x, (const long*)h_meta, k, mode, w_ms, out, ST);
}
int gg_gan_loss_bwd(const void* const* h_x, void* const* h_dx, const int64_t* h_meta, int k, int mode, float w_ms,
                    const float* gout, gg_stream_t stream) {
  return ggi_gan_loss_bwd(h_x, h_dx, (const long*)h_meta, k, mode, w_ms, gout, ST);
}
}
