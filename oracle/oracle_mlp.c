/* TEST INFRASTRUCTURE (oracle) -- see oracle.h.  Plain C restatement of
 *   - the .onnx parse the reference delegates to Ort::Session (onnx_actor.cpp:16)
 *   - the Gemm/Elu chain executed by Session::Run            (onnx_actor.cpp:47)
 * following the public ONNX opset-17 operator definitions.
 * A7 parity unpinned: ONNX Runtime 1.20.1 is not available in this image.
 */
#include "oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ wire reader */
typedef struct { const uint8_t* p; const uint8_t* end; int ok; } rd_t;

static uint64_t rd_varint(rd_t* r) {
  uint64_t v = 0; int shift = 0;
  while (r->p < r->end) {
    uint8_t b = *r->p++;
    v |= (uint64_t)(b & 0x7F) << shift;
    if (!(b & 0x80)) return v;
    shift += 7;
    if (shift > 63) break;
  }
  r->ok = 0; return 0;
}

/* returns 0 at end-of-message; fills field number, wire type and a sub-reader / scalar */
static int rd_field(rd_t* r, int* fno, int* wt, uint64_t* scalar, rd_t* sub) {
  if (!r->ok || r->p >= r->end) return 0;
  uint64_t key = rd_varint(r);
  if (!r->ok) return 0;
  *fno = (int)(key >> 3); *wt = (int)(key & 7);
  sub->ok = 1; sub->p = sub->end = r->p;
  switch (*wt) {
    case 0: *scalar = rd_varint(r); break;
    case 1: if (r->end - r->p < 8) { r->ok = 0; return 0; } sub->p = r->p; sub->end = r->p + 8; r->p += 8; break;
    case 5: if (r->end - r->p < 4) { r->ok = 0; return 0; } sub->p = r->p; sub->end = r->p + 4; r->p += 4; break;
    case 2: {
      uint64_t n = rd_varint(r);
      if (!r->ok || (uint64_t)(r->end - r->p) < n) { r->ok = 0; return 0; }
      sub->p = r->p; sub->end = r->p + n; r->p += n; break;
    }
    default: r->ok = 0; return 0;
  }
  return r->ok;
}

static void rd_str(const rd_t* s, char* dst, int cap) {
  int n = (int)(s->end - s->p);
  if (n > cap - 1) n = cap - 1;
  memcpy(dst, s->p, (size_t)n); dst[n] = 0;
}

#define MAXT 64
typedef struct { char name[96]; int64_t dims[4]; int nd; float* data; int64_t count; } tens_t;
typedef struct { char op[16]; char name[96]; char in[3][96]; int nin; char out[96];
                 float alpha, beta; int transA, transB; int has_alpha; } node_t;

static int parse_tensor(rd_t t, tens_t* o, char* err, int el) {
  int fno, wt; uint64_t sc; rd_t sub;
  int dtype = 0; const uint8_t* raw = NULL; int64_t rawn = -1;
  float* fdata = NULL; int64_t nf = 0, capf = 0;
  memset(o, 0, sizeof(*o));
  while (rd_field(&t, &fno, &wt, &sc, &sub)) {
    if (fno == 1) {
      if (wt == 0) { if (o->nd < 4) o->dims[o->nd++] = (int64_t)sc; }
      else { while (sub.p < sub.end && sub.ok) { uint64_t d = rd_varint(&sub); if (o->nd < 4) o->dims[o->nd++] = (int64_t)d; } }
    } else if (fno == 2) dtype = (int)sc;
    else if (fno == 4) {
      int64_t n = (wt == 2) ? (sub.end - sub.p) / 4 : 1;
      if (nf + n > capf) { capf = (nf + n) * 2; fdata = (float*)realloc(fdata, (size_t)capf * 4); }
      memcpy(fdata + nf, sub.p, (size_t)n * 4); nf += n;
    } else if (fno == 8) rd_str(&sub, o->name, sizeof(o->name));
    else if (fno == 9) { raw = sub.p; rawn = sub.end - sub.p; }
    else if (fno == 13 || (fno == 14 && sc != 0)) { snprintf(err, el, "initializer '%s': external data unsupported", o->name); free(fdata); return -1; }
  }
  if (!t.ok) { snprintf(err, el, "malformed TensorProto"); free(fdata); return -1; }
  if (dtype != 1) { snprintf(err, el, "initializer '%s': only FLOAT supported (got %d)", o->name, dtype); free(fdata); return -1; }
  o->count = 1; for (int i = 0; i < o->nd; i++) o->count *= o->dims[i];
  o->data = (float*)malloc((size_t)o->count * 4 + 4);
  if (raw) {
    if (rawn != o->count * 4) { snprintf(err, el, "initializer '%s': raw_data size mismatch", o->name); free(fdata); return -1; }
    memcpy(o->data, raw, (size_t)rawn);   /* little-endian host assumed */
  } else {
    if (nf != o->count) { snprintf(err, el, "initializer '%s': float_data size mismatch", o->name); free(fdata); return -1; }
    memcpy(o->data, fdata, (size_t)nf * 4);
  }
  free(fdata);
  return 0;
}

static int parse_node(rd_t n, node_t* o, char* err, int el) {
  int fno, wt; uint64_t sc; rd_t sub;
  memset(o, 0, sizeof(*o)); o->alpha = 1.f; o->beta = 1.f;
  while (rd_field(&n, &fno, &wt, &sc, &sub)) {
    if (fno == 1) { if (o->nin < 3) rd_str(&sub, o->in[o->nin], 96); o->nin++; }
    else if (fno == 2) rd_str(&sub, o->out, 96);
    else if (fno == 3) rd_str(&sub, o->name, 96);
    else if (fno == 4) rd_str(&sub, o->op, 16);
    else if (fno == 7) { if (sub.end > sub.p) { snprintf(err, el, "node '%s': non-default domain", o->name); return -1; } }
    else if (fno == 5) {
      char an[32] = {0}; float f = 0; int64_t iv = 0; int f2, w2; uint64_t s2; rd_t q;
      while (rd_field(&sub, &f2, &w2, &s2, &q)) {
        if (f2 == 1) rd_str(&q, an, sizeof(an));
        else if (f2 == 2) memcpy(&f, q.p, 4);
        else if (f2 == 3) iv = (int64_t)s2;
      }
      if (!strcmp(an, "alpha")) { o->alpha = f; o->has_alpha = 1; }
      else if (!strcmp(an, "beta")) o->beta = f;
      else if (!strcmp(an, "transA")) o->transA = (int)iv;
      else if (!strcmp(an, "transB")) o->transB = (int)iv;
    }
  }
  if (!n.ok) { snprintf(err, el, "malformed NodeProto"); return -1; }
  return 0;
}

static void parse_value_info(rd_t v, char* name, int64_t* shape, int* rank) {
  int fno, wt; uint64_t sc; rd_t sub, a, b, c, d, e;
  *rank = 0;
  while (rd_field(&v, &fno, &wt, &sc, &sub)) {
    if (fno == 1) rd_str(&sub, name, 64);
    else if (fno == 2) {
      a = sub;
      while (rd_field(&a, &fno, &wt, &sc, &b)) if (fno == 1) {           /* tensor_type */
        while (rd_field(&b, &fno, &wt, &sc, &c)) if (fno == 2) {         /* shape */
          while (rd_field(&c, &fno, &wt, &sc, &d)) if (fno == 1) {       /* dim */
            int64_t val = -1;
            while (rd_field(&d, &fno, &wt, &sc, &e)) if (fno == 1) val = (int64_t)sc;
            if (*rank < 4) shape[(*rank)++] = val;                        /* -1 = symbolic */
          }
        }
      }
    }
  }
}

static tens_t* find_tensor(tens_t* ts, int n, const char* name) {
  for (int i = 0; i < n; i++) if (!strcmp(ts[i].name, name)) return &ts[i];
  return NULL;
}

int orc_load(const char* path, orc_model** out, char* err, int el) {
  FILE* f = fopen(path, "rb");
  if (!f) { snprintf(err, el, "cannot open '%s'", path); return -1; }
  fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
  uint8_t* buf = (uint8_t*)malloc((size_t)sz + 1);
  if (fread(buf, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); free(buf); snprintf(err, el, "short read"); return -1; }
  fclose(f);
  rd_t top = { buf, buf + sz, 1 }, g = { 0, 0, 0 }, sub;
  int fno, wt; uint64_t sc;
  while (rd_field(&top, &fno, &wt, &sc, &sub)) if (fno == 7 && wt == 2) g = sub;
  if (!top.ok || !g.ok) { free(buf); snprintf(err, el, "no GraphProto / malformed model"); return -1; }

  static const int MAXN = 2 * ORC_MAX_LAYERS;
  tens_t* ts = (tens_t*)calloc(MAXT, sizeof(tens_t)); int nt = 0;
  node_t* ns = (node_t*)calloc((size_t)MAXN, sizeof(node_t)); int nn = 0;
  char in_names[MAXT][64]; int64_t in_shapes[MAXT][4]; int in_ranks[MAXT]; int nin = 0;
  char out_name[64] = {0}; int64_t out_shape[4]; int out_rank = 0; int nout = 0;
  int rc = 0;
  while (rc == 0 && rd_field(&g, &fno, &wt, &sc, &sub)) {
    if (fno == 1) { if (nn >= MAXN) { snprintf(err, el, "too many nodes"); rc = -1; break; } rc = parse_node(sub, &ns[nn++], err, el); }
    else if (fno == 5) { if (nt >= MAXT) { snprintf(err, el, "too many initializers"); rc = -1; break; } rc = parse_tensor(sub, &ts[nt++], err, el); }
    else if (fno == 11) { if (nin < MAXT) { parse_value_info(sub, in_names[nin], in_shapes[nin], &in_ranks[nin]); nin++; } }
    else if (fno == 12) { parse_value_info(sub, out_name, out_shape, &out_rank); nout++; }
  }
  if (rc == 0 && !g.ok) { snprintf(err, el, "malformed GraphProto"); rc = -1; }
  orc_model* m = (orc_model*)calloc(1, sizeof(orc_model));
  if (rc == 0) {
    int gi = -1, ngi = 0;
    for (int i = 0; i < nin; i++) if (!find_tensor(ts, nt, in_names[i])) { gi = i; ngi++; }
    if (ngi != 1 || nout != 1) { snprintf(err, el, "expected exactly one graph input and one output"); rc = -1; }
    else {
      char cur[96]; snprintf(cur, sizeof(cur), "%s", in_names[gi]);
      snprintf(m->input_name, 64, "%s", in_names[gi]); snprintf(m->output_name, 64, "%s", out_name);
      m->input_rank = in_ranks[gi]; memcpy(m->input_shape, in_shapes[gi], sizeof(in_shapes[gi]));
      m->output_rank = out_rank; memcpy(m->output_shape, out_shape, sizeof(out_shape));
      for (int i = 0; i < nn && rc == 0; i++) {
        node_t* n = &ns[i];
        if (!strcmp(n->op, "Gemm")) {
          tens_t *W = find_tensor(ts, nt, n->in[1]), *Bv = find_tensor(ts, nt, n->in[2]);
          if (n->nin != 3 || strcmp(n->in[0], cur) || !W || !Bv || W->nd != 2) { snprintf(err, el, "node '%s': Gemm not chained / missing initializer", n->name); rc = -1; break; }
          if (n->alpha != 1.f || n->beta != 1.f || n->transA) { snprintf(err, el, "node '%s': only alpha=beta=1, transA=0 supported", n->name); rc = -1; break; }
          if (m->n_layers >= ORC_MAX_LAYERS) { snprintf(err, el, "too many layers"); rc = -1; break; }
          int L = m->n_layers++;
          int o_ = (int)(n->transB ? W->dims[0] : W->dims[1]), i_ = (int)(n->transB ? W->dims[1] : W->dims[0]);
          if (Bv->count != o_ || (L > 0 && m->dims[L] != i_)) { snprintf(err, el, "node '%s': shape mismatch", n->name); rc = -1; break; }
          m->dims[L] = i_; m->dims[L + 1] = o_;
          m->w[L] = (float*)malloc((size_t)o_ * i_ * 4); m->wt[L] = (float*)malloc((size_t)o_ * i_ * 4); m->b[L] = (float*)malloc((size_t)o_ * 4);
          for (int a = 0; a < o_; a++) for (int k = 0; k < i_; k++) {
            float v = n->transB ? W->data[(size_t)a * i_ + k] : W->data[(size_t)k * o_ + a];
            m->w[L][(size_t)a * i_ + k] = v; m->wt[L][(size_t)k * o_ + a] = v;
          }
          memcpy(m->b[L], Bv->data, (size_t)o_ * 4);
          snprintf(cur, sizeof(cur), "%s", n->out);
        } else if (!strcmp(n->op, "Elu")) {
          int L = m->n_layers - 1;
          if (L < 0 || strcmp(n->in[0], cur) || m->has_elu[L]) { snprintf(err, el, "node '%s': Elu must follow a Gemm", n->name); rc = -1; break; }
          m->has_elu[L] = 1; m->elu_alpha[L] = n->has_alpha ? n->alpha : 1.f;
          snprintf(cur, sizeof(cur), "%s", n->out);
        } else { snprintf(err, el, "node '%s': unsupported op_type '%s'", n->name, n->op); rc = -1; }
      }
      if (rc == 0 && (m->n_layers == 0 || strcmp(cur, out_name))) { snprintf(err, el, "graph output not produced by the Gemm/Elu chain"); rc = -1; }
    }
  }
  for (int i = 0; i < nt; i++) free(ts[i].data);
  free(ts); free(ns); free(buf);
  if (rc != 0) { orc_free(m); return rc; }
  *out = m; return 0;
}

void orc_free(orc_model* m) {
  if (!m) return;
  for (int i = 0; i < ORC_MAX_LAYERS; i++) { free(m->w[i]); free(m->wt[i]); free(m->b[i]); }
  free(m);
}

/* ------------------------------------------------------------------ forward */
#define ORC_MAXW 2048

/* Gemm-13: Y = A*B^T + C ; k ascending, one accumulator per output (axpy form
 * so the inner loop vectorises over outputs without re-association). */
__attribute__((target_clones("avx512f", "avx2", "default")))
static void layer_f32(const float* __restrict wt, const float* __restrict b, int in, int out,
                      const float* __restrict x, float* __restrict y) {
  for (int o = 0; o < out; o++) y[o] = b[o];
  for (int k = 0; k < in; k++) {
    const float xk = x[k]; const float* __restrict wr = wt + (size_t)k * out;
    for (int o = 0; o < out; o++) y[o] += xk * wr[o];
  }
}

/* Elu-6: x < 0 ? alpha*(exp(x)-1) : x   (NaN and -0.0 pass through) */
static void elu_f32(float* y, int n, float alpha) {
  for (int o = 0; o < n; o++) { float v = y[o]; if (v < 0.f) y[o] = alpha * (expf(v) - 1.f); }
}

void orc_forward_f32(const orc_model* m, const float* x, float* y) {
  float a[ORC_MAXW], c[ORC_MAXW]; const float* cur = x; float* nxt = a;
  for (int L = 0; L < m->n_layers; L++) {
    float* dst = (L == m->n_layers - 1) ? y : nxt;
    layer_f32(m->wt[L], m->b[L], m->dims[L], m->dims[L + 1], cur, dst);
    if (m->has_elu[L]) elu_f32(dst, m->dims[L + 1], m->elu_alpha[L]);
    cur = dst; nxt = (nxt == a) ? c : a;
  }
}

void orc_forward_f64(const orc_model* m, const float* x, double* y) {
  double a[ORC_MAXW], c[ORC_MAXW]; double* cur = a; double* nxt = c;
  for (int k = 0; k < m->dims[0]; k++) cur[k] = (double)x[k];
  for (int L = 0; L < m->n_layers; L++) {
    int in = m->dims[L], out = m->dims[L + 1];
    double* dst = (L == m->n_layers - 1) ? y : nxt;
    for (int o = 0; o < out; o++) {
      double acc = 0.0; const float* wr = m->w[L] + (size_t)o * in;
      for (int k = 0; k < in; k++) acc += cur[k] * (double)wr[k];
      acc += (double)m->b[L][o];
      if (m->has_elu[L] && acc < 0.0) acc = (double)m->elu_alpha[L] * (exp(acc) - 1.0);
      dst[o] = acc;
    }
    double* t = cur; cur = dst; nxt = t;
  }
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void orc_forward_rows_f32(const orc_model* m, const float* X, float* Y, int64_t B, int threads) {
  const int in = m->dims[0], out = m->dims[m->n_layers];
  if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(static)
  for (int64_t r = 0; r < B; r++) orc_forward_f32(m, X + r * in, Y + r * out);
}

void orc_forward_rows_f64(const orc_model* m, const float* X, double* Y, int64_t B, int threads) {
  const int in = m->dims[0], out = m->dims[m->n_layers];
  if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(static)
  for (int64_t r = 0; r < B; r++) orc_forward_f64(m, X + r * in, Y + r * out);
}

/* ---- row-blocked: RB rows share each streamed weight row (better L1/L2 reuse) ---- */
#define RB 8
__attribute__((target_clones("avx512f", "avx2", "default")))
static void layer_blk_f32(const float* __restrict wt, const float* __restrict b, int in, int out,
                          const float* __restrict x /*[RB][in]*/, float* __restrict y /*[RB][out]*/,
                          int has_elu, float alpha) {
  for (int o0 = 0; o0 < out; o0 += 64) {
    const int on = (out - o0 < 64) ? out - o0 : 64;
    float acc[RB][64];
    for (int r = 0; r < RB; r++) for (int o = 0; o < on; o++) acc[r][o] = b[o0 + o];
    for (int k = 0; k < in; k++) {
      const float* __restrict wr = wt + (size_t)k * out + o0;
      for (int r = 0; r < RB; r++) { const float xk = x[r * in + k]; for (int o = 0; o < on; o++) acc[r][o] += xk * wr[o]; }
    }
    for (int r = 0; r < RB; r++) for (int o = 0; o < on; o++) {
      float v = acc[r][o];
      if (has_elu && v < 0.f) v = alpha * (expf(v) - 1.f);
      y[r * out + o0 + o] = v;
    }
  }
}

void orc_forward_blocked_f32(const orc_model* m, const float* X, float* Y, int64_t B, int threads) {
  const int in = m->dims[0], out = m->dims[m->n_layers];
  const int64_t nblk = B / RB;
  if (threads < 1) threads = 1;
#pragma omp parallel num_threads(threads)
  {
    float* a = (float*)malloc(sizeof(float) * RB * ORC_MAXW);
    float* c = (float*)malloc(sizeof(float) * RB * ORC_MAXW);
#pragma omp for schedule(static)
    for (int64_t blk = 0; blk < nblk; blk++) {
      const float* cur = X + blk * RB * in; float* nxt = a;
      for (int L = 0; L < m->n_layers; L++) {
        float* dst = (L == m->n_layers - 1) ? (Y + blk * RB * out) : nxt;
        layer_blk_f32(m->wt[L], m->b[L], m->dims[L], m->dims[L + 1], cur, dst, m->has_elu[L], m->elu_alpha[L]);
        cur = dst; nxt = (nxt == a) ? c : a;
      }
    }
    free(a); free(c);
  }
  for (int64_t r = nblk * RB; r < B; r++) orc_forward_f32(m, X + r * in, Y + r * out);
}
