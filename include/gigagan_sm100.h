/* libgigagan_sm100.so — C ABI of the B200 (sm_100a) GigaGAN training hot path.
 *
 * Conventions (SURVEY.md 8b "lower boundary"):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named `h_*`;
 *   - the library never allocates and never synchronises; the caller owns every buffer (outputs and
 *     workspaces) and passes the stream to launch on;
 *   - returns 0 on success, negative on error; gg_last_error() gives the thread-local message;
 *   - dtype: 0 = fp32 (FFMA kernels, 1e-5 parity path), 1 = bf16 storage / fp32 accumulate (tcgen05 where the
 *     shape is dense enough, FFMA otherwise).  Small statistic tensors are always fp32.
 *   - activations are NHWC (pixel rows x channels); conv weights are "kernel layout" [Cout][KH][KW][Cin].
 *
 * Each entry point names the reference interface it replaces; paths are relative to
 * lucidrains/gigagan-pytorch @ 0806433f, gigagan_pytorch/.
 */
#ifndef GIGAGAN_SM100_H
#define GIGAGAN_SM100_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef void* gg_stream_t; /* cudaStream_t */

const char* gg_last_error(void);
int gg_version(void);
/* 1 when this build contains the tcgen05/TMA kernels (always, for sm_100a). */
int gg_has_tcgen05(void);
/* bit 0: route every convolution to the FFMA kernels (testing the tcgen05 path against them);
 * bit 1: disable the thin-layer ring kernels (conv_thin_tc.cu) so those shapes take the generic tcgen05 kernels;
 * bit 2: disable only the thin-layer weight-gradient kernel.
 * bit 3: fused attention on the first-generation kernels (8 softmax warps, chunked TMEM reads; attn_tc.cu);
 * bit 4: second-generation attention kernels (attn_tc2.cu) with 8 softmax warps in every kernel, bit 5: with 16 in every
 *        kernel (default: forward 8, backward 16; A/B measurements).
 * bit 6: keep the two-pass forward for the shared-QK L2 attention (default: single pass, the row maximum of
 *        -|q_i - k_j|^2 is the diagonal).
 * bit 7: attention forward with ONE CTA per SM (default: two CTAs per SM, single S / V / P buffers, 256 TMEM columns
 *        each; bits 4 and 5 also select one-CTA forwards).
 * Returns old flags. */
int gg_set_flags(int flags);

/* ---- convolution family (replaces F.conv2d / nn.Conv2d: gigagan_pytorch.py:402-409 grouped per-sample conv of
 * AdaptiveConv2DMod; :1608-1620, :292, :1454-1470, :1656 discriminator convs; unet_upsampler.py:82-160).
 * x [N,H,W,Cin], w [G][Cout][KH][KW][Cin] (G = N when per_sample_w else 1), y [N,OH,OW,Cout].
 * y = (act(conv + bias) + res) * gain ; act 0 none / 1 leaky-relu(0.2); bias fp32 or NULL; res like y or NULL. */
int gg_conv2d_fprop(const void* x, const void* w, const float* bias, const void* res, void* y,
                    int N, int H, int W, int Cin, int OH, int OW, int Cout, int KH, int KW, int stride, int pad,
                    int per_sample_w, int act, float gain, int dtype, gg_stream_t stream);
/* same, but output pixel (n,oy,ox) channel c is written at y[y_off + n*y_sn + oy*y_sh + ox*y_sw + c] (elements):
 * lets one 1x1 launch per filter tap scatter the data gradient of a stride-2 convolution (Downsample :289-293 and
 * the stride-2 residual 1x1 :1612) straight into the interleaved input-gradient tensor. */
int gg_conv2d_fprop_strided(const void* x, const void* w, const float* bias, void* y, int N, int H, int W, int Cin,
                            int OH, int OW, int Cout, int KH, int KW, int stride, int pad, int per_sample_w, int act,
                            float gain, int64_t y_off, int64_t y_sn, int64_t y_sh, int64_t y_sw, int dtype,
                            gg_stream_t stream);
/* dx [N,H,W,Cin] = conv-transpose(dy [N,OH,OW,Cout], w)  (autograd of the calls above, data gradient) */
int gg_conv2d_dgrad(const void* dy, const void* w, void* dx,
                    int N, int H, int W, int Cin, int OH, int OW, int Cout, int KH, int KW, int stride, int pad,
                    int per_sample_w, int dtype, gg_stream_t stream);
/* dw [G][Cout][KH][KW][Cin] fp32 (overwritten) = sum over pixels of dy (x) x  (weight gradient) */
int gg_conv2d_wgrad(const void* x, const void* dy, float* dw,
                    int N, int H, int W, int Cin, int OH, int OW, int Cout, int KH, int KW, int stride, int pad,
                    int per_sample_w, int dtype, gg_stream_t stream);

/* ---- strided batched GEMM (replaces torch.einsum/bmm at gigagan_pytorch.py:574,:579,:590,:643,:651 and
 * nn.Linear / F.linear at :302-304,:887,:1121,:1658).  C[z1,z2][m,n] = alpha * sum_k A[m,k] B[k,n] (+ bias[n]).
 * h_sa/h_sb = {batch1 stride, batch2 stride, row stride, col stride} (elements, host arrays); h_sc = {s1,s2,row}. */
int gg_bmm(const void* A, const void* B, const float* bias, void* C, int b1, int b2, int M, int N, int K,
           const int64_t* h_sa, const int64_t* h_sb, const int64_t* h_sc, float alpha, int dtype, gg_stream_t stream);

/* ---- pointwise maps with explicit derivative levels (replaces nn.LeakyReLU :109, F.relu :163, nn.GELU :738,
 * nn.SiLU/nn.Sigmoid :303-305, F.normalize's 1/max(||x||,eps) :231).  kind: 0 lrelu 1 relu 2 gelu 3 silu
 * 4 sigmoid 5 invnorm 6 rsqrt(max(x,1e-8)) (demodulation :399).  level 0: out=f(x); 1: out=a*f'(x); 2: out=a*b*f''(x). */
int gg_pw_unary(int kind, int level, const void* x, const void* a, const void* b, void* out, int64_t n, int dtype,
                gg_stream_t stream);
int gg_pw_mul(const void* a, const void* b, void* out, int64_t n, int dtype, gg_stream_t stream);
int gg_pw_axpby(float alpha, const void* x, float beta, const void* y /*nullable*/, void* out, int64_t n, int dtype,
                gg_stream_t stream);
/* x viewed [R,C]; s fp32.  mode 0: s[r]; mode 1: s[((r/P)%Ns)*C+c] (per-sample-channel, scale-major repeat as in
 * gigagan_pytorch.py:1766).  op 0 multiply, 1 add.  (SqueezeExcite scaling :1218,:1767; RMSNorm gamma :232) */
int gg_pw_bcast(const void* x, const float* s, void* out, int64_t R, int C, int P, int Ns, int mode, int op, int dtype,
                gg_stream_t stream);
/* out[r] = sum_c a*b (b nullable) ; out[n',c] = sum over samples n==n' (mod Ns) and their P rows of a*b */
int gg_red_rowdot(const void* a, const void* b, float* out, int64_t R, int C, int dtype, gg_stream_t stream);
int gg_red_dot_sc(const void* a, const void* b, float* out, int64_t R, int C, int P, int Ns, int dtype,
                  gg_stream_t stream);
/* same reduction ADDED to out (a running fp32 gradient, e.g. a bias's slice of the flat gradient buffer) */
int gg_red_dot_sc_acc(const void* a, const void* b, float* out, int64_t R, int C, int P, int Ns, int dtype,
                      gg_stream_t stream);
/* row softmax over the last axis of [R,C] of (s + bias); bias fp32 [Ns][C] (nullable), row r uses bias row
 * (r / P) % Ns  (gigagan_pytorch.py:584-588 with the key bias of the L2 logits, :649; attend.py:104) */
int gg_softmax_rows(const void* s, const float* bias, void* p, int64_t R, int C, int P, int Ns, int dtype,
                    gg_stream_t stream);

/* ds = p * (gp - rowsum(p * gp)): softmax backward for one [R,C] matrix (autograd of the softmax at :588) */
int gg_softmax_bwd_rows(const void* p, const void* gp, void* ds, int64_t R, int C, int dtype, gg_stream_t stream);
/* same with a second gradient of the same shape added to gp on the fly: ds = p * ((gp + gp2) - rowsum(p * (gp + gp2)));
 * gp2 may be NULL.  Used by the attention node of the gradient-penalty pass (the probabilities collect a second-order
 * gradient besides the one arriving through the value product). */
int gg_softmax_bwd_rows_add(const void* p, const void* gp, const void* gp2, void* ds, int64_t R, int C, int dtype,
                            gg_stream_t stream);

/* second-order softmax backward (double backward of :588 inside the gradient penalty), one pass:
 * d_gp = p*(G - <G,p>), d_p = G*(gp - <p,gp>) - gp*<G,p>.  Row length must be a multiple of the 16-byte vector and
 * <= 1280 (bf16) / 640 (fp32); otherwise returns an error and the caller composes it from the primitives. */
int gg_softmax_bwd2_rows(const void* p, const void* gp, const void* G, void* d_p, void* d_gp, int64_t R, int C, int dtype,
                         gg_stream_t stream);

/* ---- separable sparse resampling of NHWC maps: bilinear x2 + [1,2,1]^2/16 reflect blur (:246-261), bilinear
 * F.interpolate (:1683-1687) and their transposes.  Tap tables: iy/wy [OH][Ty], ix/wx [OW][Tx]. */
int gg_resample2d(const void* x, void* y, int N, int H, int W, int C, int OH, int OW, const int* iy, const float* wy,
                  int Ty, const int* ix, const float* wx, int Tx, int dtype, gg_stream_t stream);

/* ---- layout at the API edge: reference tensors are NCHW fp32 (images, rgbs) */
int gg_nchw_to_nhwc(const float* src, void* dst, int N, int C, int HW, int Cpad, int dtype, gg_stream_t stream);
int gg_nhwc_to_nchw(const void* src, float* dst, int N, int C, int HW, int Cpad, int dtype, gg_stream_t stream);

/* ---- generator Noise + LeakyReLU (gigagan_pytorch.py:925-940 then :1222) */
int gg_noise_act_fwd(const void* x, const float* noise, const float* wn, void* y, int64_t R, int C, int dtype,
                     gg_stream_t stream);
int gg_noise_act_bwd(const void* y, const void* gy, const float* noise, void* dx, float* dwn, int64_t R, int C,
                     int dtype, gg_stream_t stream);

/* ---- AdaptiveConv2DMod weight builder (gigagan_pytorch.py:378-400): softmax over the filter bank, (mod+1)
 * modulation, demodulation.  bank [n][O][I][KK] fp32; mod [B][I]; kmod [B][n] (NULL if n==1);
 * w [B][Opad][KK][I] (rows O..Opad-1 are left untouched: pre-zeroed padding so Cout is a multiple of 16);
 * attn [B][n], dinv [B][O] saved for backward. */
/* mod_ld / kmod_ld: row strides (elements) of mod / kmod, so that column slices of the style projection
 * (gigagan_pytorch.py:1196 `conv_mods = ...split(...)`) are read in place. */
int gg_adaconv_weights_fwd(const float* bank, const float* mod, const float* kmod, void* w, float* attn, float* dinv,
                           int B, int n, int O, int I, int KK, int demod, float eps, int Opad, int64_t mod_ld,
                           int64_t kmod_ld, int dtype, gg_stream_t stream);
/* gattn_ws: workspace of B*n + B*O floats; dmod [B][I] and dkmod [B][n] are dense outputs */
int gg_adaconv_weights_bwd(const float* bank, const float* mod, const float* attn, const float* dinv, const float* gw,
                           float* dbank, float* dmod, float* dkmod, float* gattn_ws, int B, int n, int O, int I, int KK,
                           int demod, float eps, int Opad, int64_t mod_ld, gg_stream_t stream);

/* ---- AdaptiveConv2DMod in "shared bank" form for the low-resolution layers (4x4, 8x8), same function as
 * gigagan_pytorch.py:378-409:   y_b = dinv_b (.) sum_n attn_bn conv(x_b * (mod_b + 1), W_n)
 * (the convolution is linear in the filter, so modulation moves to the input and demodulation to the output): dense
 * convolutions over the SHARED bank whose 128-row tiles span images, instead of B private 512x512x9 filters.
 *   prep:        attn = softmax(kmod), dinv[b][o] = rsqrt(max(sum_{i,kk} ((mod+1) sum_n attn W_n)^2, eps)), xs = x*(mod+1)
 *   combine_fwd: y[b,p,o] = dinv[b,o] * sum_n attn[b,n] * ycat[b,p,n*O+o]   (ycat = conv(xs, [W_0;..;W_{n-1}]))
 *   combine_bwd: gyn[n][b,p,o] = gy*dinv*attn_n;  gdinv[b,o] = sum_p gy * sum_n attn_n ycat_n;
 *                gattn_ws[b*n+j] = sum_{p,o} gy*dinv*ycat_j   (gattn_ws: B*n floats, overwritten)
 *   bwd_stats:   the demodulation chain's gradients: dbank = dw_add (kernel-layout [n][O][KK][I] fp32 weight gradients of
 *                the n convolutions, transposed to the bank's layout) + chain term; dmod [B][I] (overwritten);
 *                dkmod [B][n] = softmax backward of (gattn_ws + chain term); gattn_ws must hold combine_bwd's result and
 *                B*O extra floats
 *   bwd_x:       gx = gxs * (mod+1);  dmod[b,i] += sum_p gxs * x                                    */
int gg_sbank_prep(const float* bank, const float* mod, const float* kmod, const void* x, void* xs, float* attn, float* dinv,
                  int B, int n, int O, int I, int KK, int HW, int demod, float eps, int64_t mod_ld, int64_t kmod_ld,
                  int dtype, gg_stream_t stream);
int gg_sbank_combine_fwd(const void* ycat, const float* attn, const float* dinv, void* y, int B, int HW, int n, int O,
                         int dtype, gg_stream_t stream);
int gg_sbank_combine_bwd(const void* gy, const void* ycat, const float* attn, const float* dinv, void* gyn, float* gdinv,
                         float* gattn_ws, int B, int HW, int n, int O, int dtype, gg_stream_t stream);
int gg_sbank_bwd_stats(const float* bank, const float* mod, const float* attn, const float* dinv, const float* gdinv,
                       const float* dw_add, float* dbank, float* dmod, float* dkmod, float* gattn_ws, int B, int n, int O,
                       int I, int KK, float eps, int64_t mod_ld, gg_stream_t stream);
int gg_sbank_bwd_x(const void* gxs, const void* x, const float* mod, void* gx, float* dmod, int B, int HW, int I,
                   int64_t mod_ld, int dtype, gg_stream_t stream);

/* ---- fused attention (gigagan_pytorch.py:562-592 with null key/value and L2-distance logits; attend.py:64-110)
 * q,k,v,o: [B, n, heads, d] rows with the given row strides (elements); null_kv [2][heads][d] fp32 or NULL.
 * mode 0 dot-product, 1 shared-QK L2 distance.  lse [B*heads][nq] fp32 saved for backward (opaque: natural-log
 * units from the FFMA kernels, log2 units from the tcgen05 kernels; fwd and bwd of one call pair agree).
 * bf16 with dim_head 64 and tokens %% 128 == 0 runs on tcgen05 (QK^T, PV, and all five backward products on the
 * tensor cores, TMEM accumulators, softmax out of TMEM); ksq_ws [B*heads*nk] fp32 workspace for the L2 form;
 * go/dq/dk/dv of the backward are dense (B, n, heads*d). */
int gg_attn_fwd(const void* q, const void* k, const void* v, const float* null_kv, void* o, float* lse, float* ksq_ws,
                int B, int heads, int nq, int nk, int d, int64_t q_rs, int64_t k_rs, int64_t v_rs, int64_t o_rs,
                float scale, int mode, int dtype, gg_stream_t stream);
int gg_attn_bwd(const void* q, const void* k, const void* v, const float* null_kv, const void* o, const void* go,
                const float* lse, void* dq, void* dk, void* dv, float* dnull_kv, float* delta_ws, float* ksq_ws,
                int B, int heads, int nq, int nk, int d, int64_t q_rs, int64_t k_rs, int64_t v_rs, int64_t o_rs,
                float scale, int mode, int dtype, gg_stream_t stream);

/* ---- operand builder of the gradient-penalty attention node (shared-QK L2-distance form, bf16, dim_head 64; replaces the
 * concatenations / casts / row reductions around gigagan_pytorch.py:574-586 on that path and their autograd):
 *   fwd:  qa [n,seq,h,80] = [q, 1, 1, 0..], ka [n,Lp,h,80] = [k, hi, lo, 0..] (k_0 = null key, k_j = q_{j-1}, padding rows 0
 *         with hi = -1e30; hi + lo = -|k|^2/2), vf [n,Lp,h,64] = [null value; v; 0]
 *   bwd:  dq = dqa[:, :64] + dka[1:seq+1, :64] - dka[1:seq+1, 64] * q,  dv = dvf[1:seq+1],  dnull [2,h,64] fp32 (j = 0 rows)
 *   bwd2: the adjoint of bwd for the cotangents (wq, wv, wnull (nullable)): g_dqa, g_dka, g_dvf, g_q = -dka[.,64] * wq, g_null */
int gg_attn_augment_fwd(const void* q, const void* v, const float* null_kv, void* qa, void* ka, void* vf, int n, int seq, int Lp,
                        int heads, int d, gg_stream_t stream);
int gg_attn_augment_bwd(const void* dqa, const void* dka, const void* dvf, const void* q, const float* null_kv, void* dq, void* dv,
                        float* dnull, int n, int seq, int Lp, int heads, int d, gg_stream_t stream);
int gg_attn_augment_bwd2(const void* wq, const void* wv, const float* wnull, const void* q, const float* null_kv, const void* dka,
                         void* g_dqa, void* g_dka, void* g_dvf, void* g_q, float* g_null, int n, int seq, int Lp, int heads, int d,
                         gg_stream_t stream);

/* ---- AdamW on a flat fp32 parameter buffer (optimizer.py:10-34 as GigaGAN configures it: betas (0.5,0.9),
 * decoupled wd on ndim>=2 tensors).  chunks: int4 {offset_lo31, len, wd_flag, offset_hi}. */
int gg_adamw(float* p, const float* g, float* m, float* v, const void* chunks, int nchunks, const int* step_ptr,
             float lr, float b1, float b2, float eps, float wd, float grad_scale, gg_stream_t stream);
int gg_incr(int* p, gg_stream_t stream);

/* ---- UnetUpsampler extras (unet_upsampler.py): 2x2 max-pool of NHWC maps (:158) and its gradient; softmax over the
 * token axis of (B, n, C) maps per (sample, channel) (LinearAttention k.softmax(dim=-1), :340). */
int gg_maxpool2_fwd(const void* x, void* y, int N, int H, int W, int C, int dtype, gg_stream_t stream);
int gg_maxpool2_bwd(const void* x, const void* gy, void* gx, int N, int H, int W, int C, int dtype, gg_stream_t stream);
int gg_softmax_tokens(const void* x, void* y, int B, int n, int C, int dtype, gg_stream_t stream);
/* Debug only (tools/bench_mma_chain.py): one warp issues `iters` tcgen05.mma (M=128, N, K=16, zero operands) that cycle
 * over `nacc` TMEM accumulators; out[0] = SM cycles to issue them, out[1] = cycles until the last one completed. */
int gg_debug_mma_chain(int N, int nacc, int iters, void* out, gg_stream_t stream);

/* Debug only: point the thin-layer convolution kernel (conv_thin_tc.cu) at a device buffer of 32*64*8 uint64; CTA 0 then
 * stores clock64 at slot ((role*64 + row%64)*8 + stage) per pipeline event (tools/trace_thin.py decodes them).
 * NULL switches tracing off. */
int gg_debug_thin_trace(void* buf);

/* ---- fused ChannelRMSNorm (gigagan_pytorch.py:224-232  F.normalize(x, dim=1) * sqrt(C) * gamma) over NHWC rows:
 * y[r,c] = x[r,c] * inv[r] * s * gamma[c], inv[r] = 1/max(|x[r,:]|, 1e-12) (saved, fp32).  bwd writes gx and ADDS the
 * gamma gradient into dgamma (fp32 [C], zero-filled by the caller).  First-order only; C % 8 == 0. */
int gg_rmsnorm_fwd(const void* x, const float* gamma, void* y, float* inv, int64_t R, int C, float s, int dtype, gg_stream_t stream);
int gg_rmsnorm_bwd(const void* x, const float* gamma, const float* inv, const void* gy, void* gx, float* dgamma, int64_t R, int C,
                   float s, int dtype, gg_stream_t stream);

/* Backward of the conv epilogue "bias + LeakyReLU(0.2)" (nn.Conv2d followed by leaky_relu, gigagan_pytorch.py:1608-1620,
 * :1454-1470): out = gy * lrelu'(y) and dbias[c] = sum over rows of out (fp32; overwritten, or added to when
 * accumulate != 0: dbias is then the bias's slice of the flat gradient buffer) in one pass.
 * Returns 1 (nothing done) when C / (16-byte vector) is not a power of two <= 256: the caller then composes
 * gg_pw_unary(level 1) + gg_red_dot_sc. */
int gg_lrelu_bwd_bias(const void* y, const void* gy, void* out, float* dbias, int64_t R, int C, int accumulate, int dtype,
                      gg_stream_t stream);

/* Accumulate a kernel-layout fp32 weight gradient dw[O][KK][Ipad] (output of gg_conv2d_wgrad) into the master-layout
 * gradient buffer dst[O][I][KK] (+=): the .grad accumulation of nn.Conv2d weights (torch autograd AccumulateGrad under
 * gigagan_pytorch.py:2113 / :2207 accelerator.backward). */
int gg_wgrad_sink(const float* dw, float* dst, int O, int I, int KK, int Ipad, gg_stream_t stream);

/* ---- once-per-step re-layout of all conv weights of a model from the flat fp32 master buffer (reference layout
 * [O][I][KK]) into both kernel layouts: fwd [O][KK][Ipad] and bwd [Ipad][KK reversed][O] (replaces the per-call
 * weight permutes cuDNN does internally for nn.Conv2d / F.conv2d and their autograd).
 * entries: int64[8] {src_off, O, I, KK, Ipad, fwd_off, bwd_off, 0}; chunks: int32[4] {entry, o0, i0, TI}:
 * one block re-lays the (32 output channels) x (TI input channels) x KK tile through shared memory. */
int gg_weight_prep_multi(const float* master, const void* entries, const void* chunks, int nchunks, void* fwd, void* bwd,
                         int dtype, gg_stream_t stream);

/* ---- heads with ONE output channel (Predictor.to_logits 1x1 conv gigagan_pytorch.py:1470/:1497, Discriminator.to_logits
 * Linear :1658): y[r] = sum_c x[r,c] w[c] + bias[0] (fp32 y), rows = pixels or flattened maps.
 * bwd: dx[r,c] = gy[r] w[c] (dx NULL: skipped);  dw[c] += sum_r gy[r] x[r,c] and dbias[0] += sum_r gy[r] (fp32, ADDED:
 * pass zeroed buffers or the parameters' slices of the flat gradient buffer; NULL: skipped). */
int gg_row_linear_fwd(const void* x, const float* w, const float* bias, float* y, int64_t R, int C, int dtype,
                      gg_stream_t stream);
int gg_row_linear_bwd(const void* x, const float* w, const float* gy, void* dx, float* dw, float* dbias, int64_t R, int C,
                      int dtype, gg_stream_t stream);

/* ---- random patch selection of the auxiliary reconstruction decoder (gigagan_pytorch.py:1300-1312: rearrange into
 * patch_dim^2 patches, keep the nsel randomly chosen ones per image).  src [B][pd*hh][pd*ww][C] -> dst [B*nsel][hh][ww][C],
 * row b*nsel+s = patch sel[b*nsel+s] (= py*pd+px) of image b.  transposed != 0: the adjoint (dst [B][pd*hh][pd*ww][C] fully
 * written: selected patches get the rows of src [B*nsel][hh][ww][C], the rest zeros). */
int gg_patch_select(const void* src, void* dst, const int* sel, int B, int nsel, int pd, int hh, int ww, int C,
                    int transposed, int dtype, gg_stream_t stream);

/* ---- GAN hinge objective over all logit tensors of a pass in ONE launch (replaces discriminator_hinge_loss /
 * generator_hinge_loss gigagan_pytorch.py:159-163 and their weighted sums :2327-2347, :2538-2551).
 * h_x: host array of k <= 8 device pointers; h_meta: host int64 [k][4] = {elements, row length, split, dtype}: every
 * row holds `split` real logits followed by fake logits (the trainer stacks real and fake images in one batch).
 * mode 0 (discriminator): L_j = mean_real relu(1 + x) + mean_fake relu(1 - x);  mode 1 (generator): L_j = mean(x).
 * out[0] = L_0, out[1] = sum_{j>=1} L_j, out[2] = L_0 + w_ms * out[1].  bwd writes dL/dx_j for out[2] scaled by the
 * device scalar gout. */
int gg_gan_loss_fwd(const void* const* h_x, const int64_t* h_meta, int k, int mode, float w_ms, float* out, gg_stream_t stream);
int gg_gan_loss_bwd(const void* const* h_x, void* const* h_dx, const int64_t* h_meta, int k, int mode, float w_ms,
                    const float* gout, gg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
