/*
 * clstm.h — C ABI of the B200-native ConvLSTM encoder-forecaster hot path (libclstm.so).
 *
 * The reference (openclimatefix/satflow v0.3.36) is pure Python: its "FFI" for this path is
 * the set of torch calls made by
 *     satflow/models/layers/ConvLSTM.py:42-57   ConvLSTMCell.forward   (cat, Conv2d, split, sigmoid/tanh, c/h update)
 *     satflow/models/layers/ConvLSTM.py:59-64   ConvLSTMCell.init_hidden
 *     satflow/models/conv_lstm.py:171-203       ConvLSTM.autoencoder   (encoder loop, decoder loop, Conv3d head, Sigmoid)
 *     satflow/models/conv_lstm.py:205-228       ConvLSTM.forward
 *     autograd of the above (training_step conv_lstm.py:53-70 -> loss.backward())
 * Each entry point below names the reference lines it replaces.  INTEGRATION.md shows the
 * ctypes binding a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative CLSTM_E* code on failure;
 *     clstm_last_error() returns a human-readable message for the calling thread;
 *   - all data pointers are DEVICE pointers owned by the caller (e.g. torch tensors);
 *     tensors in the reference's layout are contiguous fp32;
 *   - the library owns only opaque plans; a plan is used by one host thread / stream at a time;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     all work is enqueued asynchronously on it, nothing synchronises the device;
 *   - one process per GPU; multi-GPU gradient exchange is the caller's NCCL all-reduce over the
 *     gradient tensors (the path shards by batch, no collective inside the library).
 *   - there is NO CPU fallback: without a compute-capability-10.x device every call fails.
 */
#ifndef CLSTM_H_
#define CLSTM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLSTM_ABI_VERSION 2

enum {
  CLSTM_OK = 0,
  CLSTM_EINVAL = -1,   /* bad argument / unsupported configuration */
  CLSTM_ECUDA = -2,    /* CUDA runtime / driver error (message has the detail) */
  CLSTM_ENODEV = -3,   /* no sm_100 device */
  CLSTM_ESTATE = -4    /* call order violated (e.g. forward before bind / set_weights) */
};

enum { CLSTM_F16 = 0, CLSTM_BF16 = 1 }; /* operand storage type; accumulation is always fp32 */

/* Shape + hyper-parameters of one rollout (ConvLSTM.__init__ conv_lstm.py:122-169 and the
 * shapes ConvLSTM.forward sees at :215). */
typedef struct clstm_config {
  int32_t batch;        /* B                                                         */
  int32_t height;       /* H                                                         */
  int32_t width;        /* W                                                         */
  int32_t in_channels;  /* C  (input_channels)                                       */
  int32_t hidden;       /* hidden_dim                                                */
  int32_t out_channels; /* out_channels of decoder_CNN                               */
  int32_t n_layers;     /* encoder cells == decoder cells; the reference fixes 2     */
  int32_t kernel_h;     /* cell kernel; the reference ConvLSTM fixes (3,3)           */
  int32_t kernel_w;
  int32_t t_in;         /* seq_len                                                   */
  int32_t t_out;        /* forecast_steps (> 0; the reference raises on 0)           */
  int32_t dtype;        /* CLSTM_F16 (default, passes the 2e-3 gradient bar) / BF16  */
  int32_t training;     /* 1: keep gates / cell states for clstm_rollout_backward    */
  float grad_scale;     /* power-of-two loss scale for fp16 gradients; 0 = automatic */
} clstm_config_t;

typedef struct clstm_plan clstm_plan_t;

/* Message for the last failure on this thread ("" if none). */
const char* clstm_last_error(void);
int clstm_abi_version(void);
/* 0 if device `ordinal` is usable (compute capability 10.x), CLSTM_ENODEV otherwise. */
int clstm_device_check(int ordinal);

/* Plan lifetime.  create validates the configuration and sizes the workspace; bind attaches
 * caller-owned device memory (>= clstm_plan_workspace_bytes, 1024-byte aligned), encodes the TMA
 * tensor maps and zero-fills the initial states (ConvLSTMCell.init_hidden, layers/ConvLSTM.py:59-64). */
int clstm_plan_create(const clstm_config_t* cfg, clstm_plan_t** out);
int clstm_plan_destroy(clstm_plan_t* plan);
size_t clstm_plan_workspace_bytes(const clstm_plan_t* plan);
int clstm_plan_bind(clstm_plan_t* plan, void* workspace, size_t bytes, void* stream);

/* Repack the reference-layout fp32 parameters into the tensor-core operand layouts.
 * params = { cell[0].conv.weight, cell[0].conv.bias, ..., cell[2L-1].conv.weight, .bias,
 *            decoder_CNN.weight, decoder_CNN.bias }   (2*(2L)+2 pointers; cells in the order
 * encoder_1..encoder_L, decoder_1..decoder_L; shapes exactly as in the reference state_dict:
 * weight (4*hid, Cin+hid, kh, kw) with rows [i|f|o|g] and columns [x channels | h channels]
 * — layers/ConvLSTM.py:34-48 — and decoder_CNN.weight (C_out, hid, 1, 3, 3)).
 * Call again whenever the parameters change (after every optimizer step). */
int clstm_plan_set_weights(clstm_plan_t* plan, const float* const* params, int n_params, void* stream);

/* ConvLSTM.forward (conv_lstm.py:205-228): x (B,T_in,C,H,W) fp32 -> y (B,C_out,T_out,H,W) fp32. */
int clstm_rollout_forward(clstm_plan_t* plan, const float* x, float* y, void* stream);

/* The same with an explicit input layout: CLSTM_X_BTHWC is the channels-last (B,T_in,H,W,C) layout that the
 * reference's datasets emit (satflow/data/datasets.py:70-106, data/datamodules.py:188-219), consumed without a
 * permute copy. */
enum { CLSTM_X_BTCHW = 0, CLSTM_X_BTHWC = 1 };
int clstm_rollout_forward_layout(clstm_plan_t* plan, const float* x, int x_layout, float* y, void* stream);

/* Backward of clstm_rollout_forward (what loss.backward() does to conv_lstm.py:171-203):
 * dy, y are (B,C_out,T_out,H,W) fp32 (y = the forward output); grads has the same order and
 * shapes as `params`; accumulate != 0 adds into grads (like autograd's .grad +=), 0 overwrites.
 * Requires cfg.training = 1 and a preceding forward on the same plan. */
int clstm_rollout_backward(clstm_plan_t* plan, const float* dy, const float* y, float* const* grads,
                           int n_grads, int accumulate, void* stream);

/* Range statistics of the last clstm_rollout_backward on this plan, written to DEVICE memory (4 floats, async):
 *   out4 = { S, 1/S, max |dlogit|, max |S * dz| }
 * S is the power-of-two loss scale of the 16-bit gradient operands (chosen on the device from max |dlogit|), dz the
 * gate pre-activation gradients of every cell step (what loss.backward() propagates through conv_lstm.py:176-196).
 * max |S * dz| = +Inf means a 16-bit operand overflowed (exploding BPTT gradients, or a non-finite dy): the gradients
 * of that backward are not usable.  The caller decides when to look (satflow_b200 checks it without synchronising). */
int clstm_plan_grad_status(clstm_plan_t* plan, float* out4, void* stream);

/* How the plan executes (decided at bind time from the shape; DESIGN.md "Small shapes"):
 *   CLSTM_INFO_PERSISTENT_CHAIN  1 if the forward chain of cell steps (conv_lstm.py:176-196) runs as ONE persistent
 *                                launch with the cell states resident in TMEM (every tile has its own CTA), else 0;
 *   CLSTM_INFO_GRAPH_ENABLED     1 if launch-bound forward / backward calls are replayed as CUDA graphs;
 *   CLSTM_INFO_GRAPH_CAPTURES / _REPLAYS   how many graphs were captured / how many calls were replays so far. */
enum { CLSTM_INFO_PERSISTENT_CHAIN = 0, CLSTM_INFO_GRAPH_ENABLED = 1, CLSTM_INFO_GRAPH_CAPTURES = 2, CLSTM_INFO_GRAPH_REPLAYS = 3 };
int clstm_plan_info(const clstm_plan_t* plan, int what, long long* out);

/* Read back a recurrent state in the reference layout (B,hid,H,W) fp32: cell in [0,2L),
 * step in [0,T_cell] where 0 is the zero initial state and T_cell the final state.
 * Either output may be NULL.  In inference plans only the last two c steps are retained. */
int clstm_plan_read_state(clstm_plan_t* plan, int cell, int step, float* h_out, float* c_out, void* stream);

/* Measurement hook: re-launches exactly one of the kernels that the rollout issues for (cell, step), on the
 * plan's own tensors, so a caller can bracket N launches with CUDA events on `stream` (bench.py roofline,
 * tools/kernel_bench.py).  CELL_FWD is idempotent; the backward kinds reuse whatever the last backward left
 * in the scratch tensors (timing only; gate-grad updates dc in place, wgrad accumulates into its partials). */
enum {
  CLSTM_KERNEL_CELL_FWD = 0,  /* fused conv + LSTM epilogue (layers/ConvLSTM.py:45-55) */
  CLSTM_KERNEL_GATE_GRAD = 1, /* pointwise gate gradient */
  CLSTM_KERNEL_DGRAD = 2,     /* data gradient GEMM */
  CLSTM_KERNEL_WGRAD = 3,     /* weight gradient GEMM */
  CLSTM_KERNEL_DGRAD_FUSED = 5 /* data gradient GEMM of (cell, step) whose epilogue also runs the gate gradient of
                                  cell-1 at the same step (the default backward schedule; needs cell >= 1) */
};
int clstm_plan_profile_kernel(clstm_plan_t* plan, int kind, int cell, int step, void* stream);

/* ---- single cell step: ConvLSTMCell.forward (layers/ConvLSTM.py:42-57) and its backward ------
 * A cell plan is a rollout plan restricted to one cell and one step; tensors use the reference
 * layout: x (B,Cin,H,W), h/c (B,hid,H,W), weight (4*hid, Cin+hid, kh, kw), bias (4*hid) or NULL. */
typedef struct clstm_cell_plan clstm_cell_plan_t;
int clstm_cell_plan_create(int batch, int height, int width, int in_channels, int hidden, int kernel_h,
                           int kernel_w, int dtype, clstm_cell_plan_t** out);
int clstm_cell_plan_destroy(clstm_cell_plan_t* plan);
size_t clstm_cell_plan_workspace_bytes(const clstm_cell_plan_t* plan);
int clstm_cell_plan_bind(clstm_cell_plan_t* plan, void* workspace, size_t bytes, void* stream);
/* The same memory as two regions, so that every forward whose backward is still pending can own its activations:
 *   saved   — written by clstm_cell_forward, read by clstm_cell_backward (packed x / h / c, gates, packed weights);
 *   scratch — reusable by any later call (loss scale, dz, fp32 dh / dx / dc, split partial sums).
 * Re-binding is a host-only operation (it re-encodes the tensor maps), so an unrolled sequence — the reference's only
 * way of using the cell, conv_lstm.py:176-196 — binds a fresh `saved` region per step and keeps it with the autograd
 * node; clstm_cell_backward must be preceded by a bind of the region its forward wrote. */
size_t clstm_cell_plan_saved_bytes(const clstm_cell_plan_t* plan);
size_t clstm_cell_plan_scratch_bytes(const clstm_cell_plan_t* plan);
int clstm_cell_plan_bind_split(clstm_cell_plan_t* plan, void* saved, size_t saved_bytes, void* scratch,
                               size_t scratch_bytes, void* stream);
int clstm_cell_forward(clstm_cell_plan_t* plan, const float* x, const float* h_cur, const float* c_cur,
                       const float* weight, const float* bias, float* h_next, float* c_next, void* stream);
/* Gradients of one cell step given dL/dh_next, dL/dc_next (either may be NULL == zero).
 * Outputs (any may be NULL): dx, dh_cur, dc_cur, dweight, dbias.  Uses the activations that the corresponding
 * clstm_cell_forward left in the currently bound `saved` region. */
int clstm_cell_backward(clstm_cell_plan_t* plan, const float* dh_next, const float* dc_next, const float* weight,
                        float* dx, float* dh_cur, float* dc_cur, float* dweight, float* dbias, void* stream);

/* ---- native-layout stepping of one cell (inference) ------------------------------------------------------------
 * clstm_cell_forward converts x / h / c from the reference's NCHW fp32 to the kernels' NHWC 16-bit / fp32 layout and
 * back on EVERY call — for a 64-channel cell at 256x256 that is more memory traffic than the cell step itself.  A
 * caller that steps the same cell repeatedly (the loop of conv_lstm.py:176-183) can keep the recurrent state inside
 * the plan instead:
 *   clstm_cell_native_load  packs whatever is non-NULL: weight (+ bias) once, x (B,Cin,H,W) per step, h / c
 *                           (B,hid,H,W) to seed the state; reset_state != 0 zero-fills the state first
 *                           (ConvLSTMCell.init_hidden, layers/ConvLSTM.py:59-64);
 *   clstm_cell_native_step  h, c <- ConvLSTMCell.forward(x, (h, c))  (layers/ConvLSTM.py:42-57): exactly one kernel;
 *   clstm_cell_native_read  unpacks the current state to the reference layout (either pointer may be NULL).
 * Results are bit-identical to chaining clstm_cell_forward.  No activations are kept: there is no backward on this path
 * (use clstm_cell_forward / clstm_cell_backward or the rollout for training). */
int clstm_cell_native_load(clstm_cell_plan_t* plan, const float* x, const float* h, const float* c, const float* weight,
                           const float* bias, int reset_state, void* stream);
int clstm_cell_native_step(clstm_cell_plan_t* plan, void* stream);
int clstm_cell_native_read(clstm_cell_plan_t* plan, float* h_out, float* c_out, void* stream);

/* ---- fused MSE loss + gradient (EncoderDecoderConvLSTM.training_step, conv_lstm.py:55-69) --------------------
 * y (B,C,T,H,W) = the rollout output, target (B,T,C,H,W) as the reference's batches; writes
 *   out[0] = mean((permute(y) - target)^2), out[1 + t] = the same mean over frame t (the reference's per-frame
 *   losses, one .item() host sync each at :66-69), dy (B,C,T,H,W) = d out[0] / d y (may be NULL).
 * partial: device scratch of B*T*C floats.  Two-stage ordered reduction: bit-reproducible. */
int clstm_mse_loss_grad(const float* y, const float* target, int batch, int channels, int t_out, int height, int width,
                        float* dy, float* partial, float* out, void* stream);

/* Number of kernels this library has launched since load (all plans, this process). */
uint64_t clstm_launch_count(void);

/* In-situ launch trace (measurement only).  clstm_trace_enable(capacity > 0) makes every subsequent launch of this
 * library record a CUDA event on the stream of the enclosing API call (capacity = maximum number of events kept;
 * 0 disables and frees them).  clstm_trace_report synchronises on the last event, writes a per-kernel table
 * (count, total ms, share, average us; intervals = time between consecutive launch completions on that stream) into
 * buf, resets the trace and returns the number of bytes written (negative: error code).  Unlike the isolated
 * timings of clstm_plan_profile_kernel these are taken under the clocks of the real step. */
int clstm_trace_enable(int capacity);
long long clstm_trace_report(char* buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* CLSTM_H_ */
