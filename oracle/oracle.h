/* TEST INFRASTRUCTURE (oracle) -- plain-C CPU restatement of the reference hot
 * path.  Linked / loaded only by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.  Never by the product.
 *
 * PARITY STATUS
 *   A7 (policy forward): "parity unpinned" -- the ONNX Runtime 1.20.1 binary that holds
 *   its arithmetic is absent and the reference holds no golden vector; pinned instead by
 *   three independent restatements (numpy fp64, this C code, torch-CPU fp64) that agree
 *   to 2e-15 and by the known-answer vectors of SURVEY.md Appendix D.
 *   A1-A6, A9, A11 and the ONNXActor wrapper: pinned to the reference's OWN
 *   controller.cpp / onnx_actor.cpp, compiled from /root/reference against stub ROS /
 *   Eigen / ORT headers (oracle/Makefile -> oracle/_ref); bit-identical over the 400-step
 *   closed-loop fixture (tests/test_ref_controller.py, tests/golden/ref_controller_trace.npz).
 *   Caveat: the stub's Quaternion arithmetic is itself a restatement of Eigen 3.4.
 */
#ifndef GO2_ORACLE_H
#define GO2_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_LAYERS 16

typedef struct orc_model {
  int n_layers;
  int dims[ORC_MAX_LAYERS + 1];    /* dims[0] = in, dims[n_layers] = out */
  float* w[ORC_MAX_LAYERS];        /* [out][in] row-major (ONNX Gemm transB=1) */
  float* wt[ORC_MAX_LAYERS];       /* [in][out] (transposed copy for the axpy form) */
  float* b[ORC_MAX_LAYERS];
  int has_elu[ORC_MAX_LAYERS];
  float elu_alpha[ORC_MAX_LAYERS];
  char input_name[64];
  char output_name[64];
  int64_t input_shape[4];
  int64_t output_shape[4];
  int input_rank, output_rank;
} orc_model;

/* reference: onnx_actor.cpp:16 (Ort::Session ctor parses the file once) */
int orc_load(const char* path, orc_model** out, char* err, int err_len);
void orc_free(orc_model* m);

/* reference: onnx_actor.cpp:47 (Session::Run), ONNX Gemm-13 / Elu-6 semantics */
void orc_forward_f32(const orc_model* m, const float* x, float* y);
void orc_forward_f64(const orc_model* m, const float* x, double* y);
/* B independent batch-1 forwards (reference semantics), rows split over threads */
void orc_forward_rows_f32(const orc_model* m, const float* X, float* Y, int64_t B, int threads);
/* row-blocked variant ("generous" CPU baseline): weights reused across RB rows */
void orc_forward_blocked_f32(const orc_model* m, const float* X, float* Y, int64_t B, int threads);
void orc_forward_rows_f64(const orc_model* m, const float* X, double* Y, int64_t B, int threads);
int orc_max_threads(void);

/* ---- controller (reference: controller.cpp:173-251, controller.hpp:45-68,92-111) ---- */
#define ORC_MAX_HIST 8
#define ORC_FRAME 49

typedef struct orc_raw_state {
  float quat[4];        /* w,x,y,z  controller.hpp:95-97 */
  float gyro[3];        /* controller.hpp:109 */
  float q[12];          /* (float)get_q()[i]  controller.cpp:189 */
  float dq[12];         /* controller.cpp:190 */
  float axes[4];        /* joy axes, 0/1/3 used  controller.cpp:176-178 */
  int16_t foot_force[4];/* unitree order  controller.hpp:100-103 */
  int32_t joy_valid;    /* joy_ && !axes.empty()  controller.cpp:173 */
  int32_t button0;      /* joy_->buttons[0]  controller.cpp:221,246 */
} orc_raw_state;

typedef struct orc_ctrl_state {
  int H;
  float vel_cmd[3];
  float g_hist[3 * ORC_MAX_HIST], w_hist[3 * ORC_MAX_HIST], cmd_hist[3 * ORC_MAX_HIST];
  float q_hist[12 * ORC_MAX_HIST], dq_hist[12 * ORC_MAX_HIST], a_hist[12 * ORC_MAX_HIST];
  uint16_t c_hist[4 * ORC_MAX_HIST];
  float action[12];
  float kp, kd;
} orc_ctrl_state;

typedef struct orc_step_out {
  float obs[ORC_FRAME * ORC_MAX_HIST];
  float action_raw[12];
  float action[12];
  double q_des[12];
  double kp, kd;
} orc_step_out;

void orc_ctrl_reset(orc_ctrl_state* s, int H);
/* obs assembly only (A1-A6); mutates s like publish() does */
void orc_ctrl_assemble(orc_ctrl_state* s, const orc_raw_state* raw, float* obs);
/* post-processing (A9, A11) on a given raw action */
void orc_ctrl_post(orc_ctrl_state* s, const orc_raw_state* raw, const float* action_raw, orc_step_out* out);
/* full step with the oracle's own forward: use_f64 selects fp64 or fp32 arithmetic for A7 */
void orc_ctrl_step(orc_ctrl_state* s, const orc_model* m, const orc_raw_state* raw, int use_f64, orc_step_out* out);
/* closed loop on one host thread, fp32 forward: `steps` publish() steps over the raw states (cycled), each timed
 * with CLOCK_MONOTONIC; ns_out[steps].  The CPU stand-in for the reference's control-loop step (SURVEY 8d config 1:
 * ONNX Runtime itself is not available). Returns the last action's first component (keeps the loop observable). */
float orc_ctrl_closed_loop_ns(const orc_model* m, int H, const orc_raw_state* raws, int n_raws, int64_t steps, uint64_t* ns_out);

#ifdef __cplusplus
}
#endif
#endif
