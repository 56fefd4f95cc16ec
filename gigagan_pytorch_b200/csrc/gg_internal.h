// Internal (C++ linkage) launchers implemented in the kernel translation units.
#pragma once
#include "gg_common.cuh"

int ggi_simt_conv_fprop(const void* x, const void* w, const float* bias, const void* res, void* y, int N, int H, int W,
                        int Cin, int OH, int OW, int Cout, int KH, int KW, int stride, int pad, int per_sample_w, int act,
                        float gain, const long* ystr, int dtype, cudaStream_t st);
int ggi_simt_conv_dgrad(const void* dy, const void* w, void* dx, int N, int H, int W, int Cin, int OH, int OW, int Cout,
                        int KH, int KW, int stride, int pad, int per_sample_w, int dtype, cudaStream_t st);
int ggi_simt_conv_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int OH, int OW, int Cout,
                        int KH, int KW, int stride, int pad, int per_sample_w, int dtype, cudaStream_t st);
int ggi_simt_bmm(const void* A, const void* B, const float* bias, void* C, int b1, int b2, int M, int N, int K,
                 const long* sa, const long* sb, const long* sc, float alpha, int dtype, cudaStream_t st);
int ggi_pw_unary(int kind, int level, const void* x, const void* a, const void* b, void* out, long n, int dtype, cudaStream_t st);
int ggi_pw_mul(const void* a, const void* b, void* out, long n, int dtype, cudaStream_t st);
int ggi_pw_axpby(float alpha, const void* x, float beta, const void* y, void* out, long n, int dtype, cudaStream_t st);
int ggi_pw_bcast(const void* x, const float* s, void* out, long R, int C, int P, int Ns, int mode, int op, int dtype, cudaStream_t st);
int ggi_red_rowdot(const void* a, const void* b, float* out, long R, int C, int dtype, cudaStream_t st);
int ggi_red_dot_sc(const void* a, const void* b, float* out, long R, int C, int P, int Ns, int dtype, cudaStream_t st);
int ggi_softmax_rows(const void* s, const float* bias, void* p, long R, int C, int P, int Ns, int dtype, cudaStream_t st);
int ggi_resample2d(const void* x, void* y, int N, int H, int W, int C, int OH, int OW, const int* iy, const float* wy,
                   int Ty, const int* ix, const float* wx, int Tx, int dtype, cudaStream_t st);
int ggi_nchw_to_nhwc(const float* src, void* dst, int N, int C, int HW, int Cp, int dtype, cudaStream_t st);
int ggi_nhwc_to_nchw(const void* src, float* dst, int N, int C, int HW, int Cp, int dtype, cudaStream_t st);
int ggi_noise_act_fwd(const void* x, const float* noise, const float* wn, void* y, long R, int C, int dtype, cudaStream_t st);
int ggi_noise_act_bwd(const void* y, const void* gy, const float* noise, void* dx, float* dwn, long R, int C, int dtype, cudaStream_t st);
int ggi_adaconv_weights_fwd(const float* bank, const float* mod, const float* kmod, void* w, float* attn, float* dinv,
                            int B, int n, int O, int I, int KK, int demod, float eps, int Opad, long ldm, long ldk, int dtype,
                            cudaStream_t st);
int ggi_adaconv_weights_bwd(const float* bank, const float* mod, const float* attn, const float* dinv, const float* gw,
                            float* dbank, float* dmod, float* dkmod, float* gattn_ws, int B, int n, int O, int I, int KK,
                            int demod, float eps, int Opad, long ldm, const float* q_ext, const float* dw_add,
                            cudaStream_t st);
int ggi_red_dot_sc_acc(const void* a, const void* b, float* out, long R, int C, int P, int Ns, int accumulate, int dtype,
                       cudaStream_t st);
int ggi_attn_fwd(const void* q, const void* k, const void* v, const float* null_kv, void* o, float* lse, int B, int heads,
                 int nq, int nk, int d, long q_rs, long k_rs, long v_rs, long o_rs, float scale, int mode, int dtype,
                 cudaStream_t st);
int ggi_attn_bwd(const void* q, const void* k, const void* v, const float* null_kv, const void* o, const void* go,
                 const float* lse, void* dq, void* dk, void* dv, float* dnull_kv, float* delta_ws, int B, int heads,
                 int nq, int nk, int d, long q_rs, long k_rs, long v_rs, long o_rs, float scale, int mode, int dtype,
                 cudaStream_t st);
int ggi_adamw(float* p, const float* g, float* m, float* v, const void* chunks, int nchunks, const int* step_ptr,
              float lr, float b1, float b2, float eps, float wd, float grad_scale, cudaStream_t st);
int ggi_incr(int* p, cudaStream_t st);

// tcgen05 path (conv_tc.cu).  Return 1 when the shape is not eligible (caller falls through to FFMA), 0 ok, <0 error.
int ggi_tc_conv_thin_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int OH, int OW, int Cout,
                           int KH, int KW, int stride, int pad, int per_sample_w, cudaStream_t st);
int ggi_rmsnorm_fwd(const void* x, const float* gamma, void* y, float* inv, long R, int C, float s, int dtype, cudaStream_t st);
int ggi_rmsnorm_bwd(const void* x, const float* gamma, const float* inv, const void* gy, void* gx, float* dgamma, long R, int C,
                    float s, int dtype, cudaStream_t st);
int ggi_debug_thin_trace(unsigned long long* buf);
int ggi_debug_mma_chain(int N, int nacc, int iters, unsigned long long* out, cudaStream_t st);
int ggi_lrelu_bwd_bias(const void* y, const void* gy, void* out, float* dbias, long R, int C, int accumulate, int dtype,
                       cudaStream_t st);
int ggi_wgrad_sink(const float* dw, float* dst, int O, int I, int KK, int Ipad, cudaStream_t st);
int ggi_tc_conv_thin(const void* x, const void* w, const float* bias, const void* res, void* y, int N, int H, int W, int Cin,
                     int OH, int OW, int Cout, int KH, int KW, int stride, int pad, int per_sample_w, int act, float gain,
                     cudaStream_t st);
int ggi_tc_conv_fprop(const void* x, const void* w, const float* bias, const void* res, void* y, int N, int H, int W,
                      int Cin, int OH, int OW, int Cout, int KH, int KW, int stride, int pad, int per_sample_w, int act,
                      float gain, const long* ystr, cudaStream_t st);
int ggi_tc_conv_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int OH, int OW, int Cout,
                      int KH, int KW, int stride, int pad, int per_sample_w, cudaStream_t st);
int ggi_tc_bmm(const void* A, const void* B, const float* bias, void* C, int b1, int b2, int M, int N, int K,
               const long* sa, const long* sb, const long* sc, float alpha, cudaStream_t st);
int ggi_softmax_bwd_rows(const void* p, const void* gp, const void* gp2, void* ds, long R, int C, int dtype, cudaStream_t st);
int ggi_weight_prep_multi(const float* master, const void* entries, const void* chunks, int nchunks, void* fwd, void* bwd,
                          int dtype, cudaStream_t st);
int ggi_tc_attn_fwd(const void* q, const void* k, const void* v, const float* null_kv, void* o, float* lse, float* ksq_ws,
                    int B, int heads, int nq, int nk, int d, long q_rs, long k_rs, long v_rs, long o_rs, float scale,
                    int mode, cudaStream_t st);
int ggi_tc_attn_bwd(const void* q, const void* k, const void* v, const float* null_kv, const void* o, const void* go,
                    const float* lse2, void* dq, void* dk, void* dv, float* dnull_kv, float* delta_ws, float* ksq_ws,
                    int B, int heads, int nq, int nk, int d, long q_rs, long k_rs, long v_rs, long o_rs, float scale,
                    int mode, cudaStream_t st);
int ggi_tc2_attn_fwd(const void* q, const void* k, const void* v, const float* null_kv, void* o, float* lse, float* ksq_ws,
                     int B, int heads, int nq, int nk, int d, long q_rs, long k_rs, long v_rs, long o_rs, float scale,
                     int mode, int nsw, cudaStream_t st);
int ggi_tc2_attn_bwd(const void* q, const void* k, const void* v, const float* null_kv, const void* o, const void* go,
                     const float* lse2, void* dq, void* dk, void* dv, float* dnull_kv, float* delta_ws, float* ksq_ws,
                     int B, int heads, int nq, int nk, int d, long q_rs, long k_rs, long v_rs, long o_rs, float scale,
                     int mode, int nsw, cudaStream_t st);
int ggi_softmax_bwd2_rows(const void* p, const void* gp, const void* G, void* d_p, void* d_gp, long R, int C, int dtype,
                          cudaStream_t st);
int ggi_maxpool2_fwd(const void* x, void* y, int N, int H, int W, int C, int dtype, cudaStream_t st);
int ggi_maxpool2_bwd(const void* x, const void* gy, void* gx, int N, int H, int W, int C, int dtype, cudaStream_t st);
int ggi_softmax_tokens(const void* x, void* y, int B, int n, int C, int dtype, cudaStream_t st);
int ggi_gan_loss_fwd(const void* const* x, const long* meta, int k, int mode, float w_ms, float* out, cudaStream_t st);
int ggi_gan_loss_bwd(const void* const* x, void* const* dx, const long* meta, int k, int mode, float w_ms, const float* gout,
                     cudaStream_t st);
