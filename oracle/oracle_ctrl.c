/* TEST INFRASTRUCTURE (oracle) -- see oracle.h.  Plain C restatement of the
 * observation assembly and action post-processing of ONNXController::publish()
 * (reference: onnx_controller/src/controller.cpp:173-251,
 *             onnx_controller/include/onnx_controller/controller.hpp:45-68,92-111).
 * Compiled with -ffp-contract=off: the reference is built for baseline x86-64
 * (no FMA), so no multiply-add may be contracted here either.
 * Pinned against the reference's own controller.cpp via oracle/_ref (see Makefile).
 */
#include "oracle.h"

#include <math.h>
#include <string.h>

/* controller.hpp:165 */
static const double Q0[12] = {0.1, -0.1, 0.1, -0.1, 0.8, 0.8, 1.0, 1.0, -1.5, -1.5, -1.5, -1.5};
/* controller.hpp:16 */
static const float ACTION_LIMIT = 1000.0f;

void orc_ctrl_reset(orc_ctrl_state* s, int H) {
  memset(s, 0, sizeof(*s));          /* controller.hpp:132-162: all histories zero */
  s->H = H;
  s->kp = 28.0f; s->kd = 0.5f;       /* controller.hpp:119-120 */
}

/* populate_buffer: std::shift_left by n, copy head to the tail (controller.hpp:45-51) */
static void push_f(float* hist, int total, const float* cur, int n) {
  memmove(hist, hist + n, (size_t)(total - n) * sizeof(float));
  memcpy(hist + total - n, cur, (size_t)n * sizeof(float));
}
static void push_u16(uint16_t* hist, int total, const uint16_t* cur, int n) {
  memmove(hist, hist + n, (size_t)(total - n) * sizeof(uint16_t));
  memcpy(hist + total - n, cur, (size_t)n * sizeof(uint16_t));
}

static void cross3(const float* a, const float* b, float* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

/* controller.cpp:182-184: gb = quaternion_.inverse() * (0,0,-1)  (Eigen semantics) */
static void gravity_body(const float* quat_wxyz, float* g) {
  const float w = quat_wxyz[0], x = quat_wxyz[1], y = quat_wxyz[2], z = quat_wxyz[3];
  const float n2 = ((x * x + y * y) + z * z) + w * w;
  float c[4];                                    /* conj / n2, or zero */
  if (n2 > 0.0f) { c[0] = w / n2; c[1] = -x / n2; c[2] = -y / n2; c[3] = -z / n2; }
  else { c[0] = c[1] = c[2] = c[3] = 0.0f; }
  const float v[3] = {0.0f, 0.0f, -1.0f};
  const float* u = &c[1];
  float uv[3], uuv[3];
  cross3(u, v, uv);
  for (int i = 0; i < 3; i++) uv[i] += uv[i];
  cross3(u, uv, uuv);
  for (int i = 0; i < 3; i++) g[i] = (v[i] + c[0] * uv[i]) + uuv[i];
}

void orc_ctrl_assemble(orc_ctrl_state* s, const orc_raw_state* raw, float* obs) {
  const int H = s->H;
  /* A1  controller.cpp:173-179 */
  if (raw->joy_valid) {
    const float a0 = raw->axes[0], a1 = raw->axes[1], a3 = raw->axes[3];
    s->vel_cmd[0] = a1;
    s->vel_cmd[1] = (float)(((double)a0 * (double)a0) * ((a0 > 0) ? 1 : -1) * 0.8);
    s->vel_cmd[2] = a3 * a1;
  }
  /* A2 */
  float g[3]; gravity_body(raw->quat, g);
  /* A3  controller.cpp:187-197: float -= double */
  float q[12], dq[12];
  for (int i = 0; i < 12; i++) { q[i] = (float)((double)raw->q[i] - Q0[i]); dq[i] = raw->dq[i]; }
  /* A4  controller.hpp:99-103 */
  uint16_t c[4];
  c[0] = raw->foot_force[1] >= 22; c[1] = raw->foot_force[0] >= 22;
  c[2] = raw->foot_force[3] >= 22; c[3] = raw->foot_force[2] >= 22;
  /* A5  controller.cpp:200-207 */
  push_f(s->g_hist, 3 * H, g, 3);
  push_f(s->w_hist, 3 * H, raw->gyro, 3);
  push_f(s->cmd_hist, 3 * H, s->vel_cmd, 3);
  push_f(s->q_hist, 12 * H, q, 12);
  push_f(s->dq_hist, 12 * H, dq, 12);
  push_f(s->a_hist, 12 * H, s->action, 12);
  push_u16(s->c_hist, 4 * H, c, 4);
  /* A6  controller.cpp:210-212: seven shift-appends fill the whole buffer */
  float* o = obs;
  memcpy(o, s->g_hist, sizeof(float) * 3 * H); o += 3 * H;
  memcpy(o, s->w_hist, sizeof(float) * 3 * H); o += 3 * H;
  memcpy(o, s->cmd_hist, sizeof(float) * 3 * H); o += 3 * H;
  memcpy(o, s->q_hist, sizeof(float) * 12 * H); o += 12 * H;
  memcpy(o, s->dq_hist, sizeof(float) * 12 * H); o += 12 * H;
  memcpy(o, s->a_hist, sizeof(float) * 12 * H); o += 12 * H;
  for (int i = 0; i < 4 * H; i++) o[i] = (float)s->c_hist[i];
}

void orc_ctrl_post(orc_ctrl_state* s, const orc_raw_state* raw, const float* action_raw, orc_step_out* out) {
  for (int i = 0; i < 12; i++) {
    float a = action_raw[i];
    out->action_raw[i] = a;
    /* A9  controller.cpp:218-223: std::clamp then multiply by (buttons[0]==0) */
    a = (a < -ACTION_LIMIT) ? -ACTION_LIMIT : (ACTION_LIMIT < a) ? ACTION_LIMIT : a;
    a *= (float)(raw->button0 == 0);
    s->action[i] = a;
    out->action[i] = a;
    /* A11 controller.cpp:244 */
    out->q_des[i] = Q0[i] + a * 0.25;
  }
  out->kp = (double)((raw->button0 == 0) ? s->kp : 5.0f);
  out->kd = (double)s->kd;
}

void orc_ctrl_step(orc_ctrl_state* s, const orc_model* m, const orc_raw_state* raw, int use_f64, orc_step_out* out) {
  float a[ORC_MAX_LAYERS * 0 + 64];
  orc_ctrl_assemble(s, raw, out->obs);
  if (use_f64) {
    double d[64]; orc_forward_f64(m, out->obs, d);
    for (int i = 0; i < 12; i++) a[i] = (float)d[i];
  } else {
    orc_forward_f32(m, out->obs, a);
  }
  orc_ctrl_post(s, raw, a, out);
}

/* reference: controller.cpp:173-251 executed `steps` times, one thread, per-step wall time */
#include <time.h>
float orc_ctrl_closed_loop_ns(const orc_model* m, int H, const orc_raw_state* raws, int n_raws, int64_t steps, uint64_t* ns_out) {
  orc_ctrl_state st;
  orc_step_out out;
  orc_ctrl_reset(&st, H);
  for (int64_t i = 0; i < steps; ++i) {
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    orc_ctrl_step(&st, m, &raws[i % n_raws], 0, &out);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (ns_out) ns_out[i] = (uint64_t)((t1.tv_sec - t0.tv_sec) * 1000000000ll + (t1.tv_nsec - t0.tv_nsec));
  }
  return out.action[0];
}
