/*
 * TEST INFRASTRUCTURE — plain-C restatement of the reference hot path (goruck/radar-ml),
 * independent of the numpy one in oracle/restate.py.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load this; the product (radar_ml_b200/) never does.
 *
 * Every function cites the reference line range it follows (paths under /root/reference), or
 * scikit-learn (SK/ = site-packages/sklearn; the reference pins scikit-learn==0.24.0,
 * requirements.txt:57) for arithmetic that lives in that dependency.
 *
 * Pinned by tests/test_oracle_c.py against the golden vectors minted from the unmodified
 * reference (tests/golden/) and against oracle/restate.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define RO_MAX_CLASSES 8

typedef struct ro_model {
  int32_t kind;               /* 0: CalibratedClassifierCV(SVC rbf), 1: CalibratedClassifierCV(SGD log) */
  int32_t n_classes, n_features, n_sv;
  double gamma;
  const double* sv;           /* [n_sv][F]          SVC.support_vectors_              */
  const double* dual_coef;    /* [C-1][n_sv]        SVC._dual_coef_ (libsvm sign)      */
  const double* rho;          /* [C(C-1)/2]         = -SVC._intercept_                 */
  const int32_t* n_support;   /* [C]                                                    */
  const double* platt_a;      /* [C] ([1] if C == 2) _SigmoidCalibration.a_            */
  const double* platt_b;
  const double* coef;         /* linear: [R][F], R = C (1 if C == 2)  SGDClassifier.coef_ */
  const double* intercept;    /* linear: [R]                                            */
} ro_model;

/* predict.py:102-107 (slice through the target voxel; numpy wraps a negative index once and
 * raises IndexError otherwise -> return -1) or the north_star axis-max projections.
 * cube [sx][sy][sz] (predict.py:91, z fastest); xz [sx][sz], yz [sy][sz], xy [sx][sy]. */
int ro_project(const float* cube, int sx, int sy, int sz, int mode, int i, int j, int k,
               float* xz, float* yz, float* xy) {
  if (mode == 1) {
    if (i < 0) i += sx;
    if (j < 0) j += sy;
    if (k < 0) k += sz;
    if (i < 0 || i >= sx || j < 0 || j >= sy || k < 0 || k >= sz) return -1;
    for (int a = 0; a < sx; ++a)
      for (int c = 0; c < sz; ++c) xz[a * sz + c] = cube[((size_t)a * sy + j) * sz + c];   /* raw[:, j, :] */
    for (int b = 0; b < sy; ++b)
      for (int c = 0; c < sz; ++c) yz[b * sz + c] = cube[((size_t)i * sy + b) * sz + c];   /* raw[i, :, :] */
    for (int a = 0; a < sx; ++a)
      for (int b = 0; b < sy; ++b) xy[a * sy + b] = cube[((size_t)a * sy + b) * sz + k];   /* raw[:, :, k] */
    return 0;
  }
  for (int a = 0; a < sx; ++a)
    for (int c = 0; c < sz; ++c) {
      float m = cube[((size_t)a * sy) * sz + c];
      for (int b = 1; b < sy; ++b) { const float v = cube[((size_t)a * sy + b) * sz + c]; if (v > m) m = v; }
      xz[a * sz + c] = m;
    }
  for (int b = 0; b < sy; ++b)
    for (int c = 0; c < sz; ++c) {
      float m = cube[((size_t)b) * sz + c];
      for (int a = 1; a < sx; ++a) { const float v = cube[((size_t)a * sy + b) * sz + c]; if (v > m) m = v; }
      yz[b * sz + c] = m;
    }
  for (int a = 0; a < sx; ++a)
    for (int b = 0; b < sy; ++b) {
      const float* row = cube + ((size_t)a * sy + b) * sz;
      float m = row[0];
      for (int c = 1; c < sz; ++c) if (row[c] > m) m = row[c];
      xy[a * sy + b] = m;
    }
  return 0;
}

/* common.py:141-149 at zoom 1.0: np.concatenate((xz, yz, xy) selected by proj_mask, axis=None),
 * then `/ RADAR_MAX` as a float32 true division when scale.  mask bit0 xz, bit1 yz, bit2 xy.
 * Returns the feature length F. */
int ro_process_samples(const float* xz, const float* yz, const float* xy, int sx, int sy, int sz,
                       unsigned mask, int scale, float* out) {
  int f = 0;
  if (mask & 1u) { memcpy(out + f, xz, sizeof(float) * (size_t)sx * sz); f += sx * sz; }
  if (mask & 2u) { memcpy(out + f, yz, sizeof(float) * (size_t)sy * sz); f += sy * sz; }
  if (mask & 4u) { memcpy(out + f, xy, sizeof(float) * (size_t)sx * sy); f += sx * sy; }
  if (scale)
    for (int e = 0; e < f; ++e) out[e] = out[e] / 255.0f;          /* common.py:31, 148 */
  return f;
}

/* scipy.special.expit as SK/calibration.py:1065 uses it */
static double ro_expit(double x) {
  if (x >= 0) return 1.0 / (1.0 + exp(-x));
  const double e = exp(x);
  return e / (1.0 + e);
}

/* decision_function of one observation: dec[C] ([1] meaningful when C == 2).
 * SVC: SK/svm/src/libsvm/svm.cpp:461-514 (RBF k_function on float64, X cast at SK/svm/_base.py:590),
 * :2864-2893 (pair sums - rho), SK/svm/_base.py:798-828 + SK/utils/multiclass.py:557-599 (OvR
 * shape: votes + sum_of_confidences / (3 (|.| + 1)), called with (dec < 0, -dec)).
 * linear: train.py:368-369 SGDClassifier.decision_function = x . coef^T + intercept. */
static void ro_decision_one(const float* x, const ro_model* m, double* kv, double* dec) {
  const int C = m->n_classes, F = m->n_features;
  if (m->kind == 1) {
    const int R = C == 2 ? 1 : C;
    for (int r = 0; r < R; ++r) {
      double s = 0.0;
      for (int f = 0; f < F; ++f) s += (double)x[f] * m->coef[(size_t)r * F + f];
      dec[r] = s + m->intercept[r];
    }
    return;
  }
  for (int s = 0; s < m->n_sv; ++s) {
    const double* sv = m->sv + (size_t)s * F;
    double sum = 0.0;
    for (int f = 0; f < F; ++f) { const double d = (double)x[f] - sv[f]; sum += d * d; }
    kv[s] = exp(-m->gamma * sum);
  }
  int start[RO_MAX_CLASSES];
  start[0] = 0;
  for (int c = 1; c < C; ++c) start[c] = start[c - 1] + m->n_support[c - 1];
  double ovo[RO_MAX_CLASSES * (RO_MAX_CLASSES - 1) / 2];
  int q = 0;
  for (int i = 0; i < C; ++i)
    for (int j = i + 1; j < C; ++j, ++q) {
      double s = 0.0;
      for (int t = 0; t < m->n_support[i]; ++t)
        s += m->dual_coef[(size_t)(j - 1) * m->n_sv + start[i] + t] * kv[start[i] + t];
      double s2 = 0.0;
      for (int t = 0; t < m->n_support[j]; ++t)
        s2 += m->dual_coef[(size_t)i * m->n_sv + start[j] + t] * kv[start[j] + t];
      ovo[q] = s + s2 - m->rho[q];
    }
  if (C == 2) { dec[0] = -ovo[0]; return; }             /* SK/svm/_base.py binary sign flip */
  double votes[RO_MAX_CLASSES] = {0}, soc[RO_MAX_CLASSES] = {0};
  q = 0;
  for (int i = 0; i < C; ++i)
    for (int j = i + 1; j < C; ++j, ++q) {
      const double conf = -ovo[q];
      soc[i] -= conf;
      soc[j] += conf;
      if (ovo[q] < 0) votes[j] += 1; else votes[i] += 1;
    }
  for (int c = 0; c < C; ++c) dec[c] = votes[c] + soc[c] / (3.0 * (fabs(soc[c]) + 1.0));
}

/* SK/calibration.py:781-850: per-class sigmoid, row normalisation (uniform when the row sums
 * to zero), values in (1, 1 + 1e-5] clipped to 1. */
static void ro_platt(const double* dec, const ro_model* m, double* proba) {
  const int C = m->n_classes;
  if (C == 2) {
    proba[1] = ro_expit(-(m->platt_a[0] * dec[0] + m->platt_b[0]));
    proba[0] = 1.0 - proba[1];
  } else {
    double den = 0.0;
    for (int c = 0; c < C; ++c) { proba[c] = ro_expit(-(m->platt_a[c] * dec[c] + m->platt_b[c])); den += proba[c]; }
    for (int c = 0; c < C; ++c) proba[c] = den != 0.0 ? proba[c] / den : 1.0 / C;
  }
  for (int c = 0; c < C; ++c) if (proba[c] > 1.0 && proba[c] <= 1.0 + 1e-5) proba[c] = 1.0;
}

/* model.predict_proba(X) (predict.py:60) for n rows of F float32 features; decision (nullable)
 * receives the OvR decision values [n][C] ([n] when C == 2). */
int ro_predict_proba(const float* X, int64_t n, const ro_model* m, double* proba, double* decision) {
  if (m->n_classes < 2 || m->n_classes > RO_MAX_CLASSES) return -1;
  const int C = m->n_classes, R = C == 2 ? 1 : C;
  int rc = 0;
#pragma omp parallel
  {
    double* kv = (double*)malloc(sizeof(double) * (size_t)(m->n_sv > 0 ? m->n_sv : 1));
    if (!kv) {
#pragma omp atomic write
      rc = -2;
    }
#pragma omp for schedule(dynamic, 8)
    for (int64_t r = 0; r < n; ++r) {
      if (!kv) continue;
      double dec[RO_MAX_CLASSES];
      ro_decision_one(X + r * m->n_features, m, kv, dec);
      if (decision) for (int c = 0; c < R; ++c) decision[r * R + c] = dec[c];
      ro_platt(dec, m, proba + r * C);
    }
    free(kv);
  }
  return rc;
}

/* predict.py:90-119 for n scans: projection -> process_samples(scale=True) -> classifier().
 * feats (nullable) [n][F] float32; proba [n][C]; label [n] = argmax (first maximum);
 * known [n] = (max proba >= min_proba), i.e. name != 'Unknown' (predict.py:65-68).
 * Returns 0, or -(1 + index) of the first scan whose slice index numpy would reject. */
int64_t ro_scan_path(const float* cubes, int64_t n, int sx, int sy, int sz, int mode,
                     const int32_t* ijk, unsigned mask, const ro_model* m, double min_proba,
                     float* feats, double* proba, int32_t* label, uint8_t* known) {
  if (m->n_classes < 2 || m->n_classes > RO_MAX_CLASSES) return -1;
  const int C = m->n_classes;
  const size_t cube_elems = (size_t)sx * sy * sz;
  int F = 0;
  if (mask & 1u) F += sx * sz;
  if (mask & 2u) F += sy * sz;
  if (mask & 4u) F += sx * sy;
  if (F != m->n_features) return -1;
  int64_t bad = 0;
#pragma omp parallel
  {
    float* xz = (float*)malloc(sizeof(float) * ((size_t)sx * sz + (size_t)sy * sz + (size_t)sx * sy + (size_t)F));
    double* kv = (double*)malloc(sizeof(double) * (size_t)(m->n_sv > 0 ? m->n_sv : 1));
    float* yz = xz + (size_t)sx * sz;
    float* xy = yz + (size_t)sy * sz;
    float* row = xy + (size_t)sx * sy;
#pragma omp for schedule(dynamic, 8)
    for (int64_t s = 0; s < n; ++s) {
      if (!xz || !kv) continue;
      const int32_t* t = ijk ? ijk + s * 3 : NULL;
      if (ro_project(cubes + s * cube_elems, sx, sy, sz, mode, t ? t[0] : 0, t ? t[1] : 0, t ? t[2] : 0,
                     xz, yz, xy) != 0) {
#pragma omp critical
        { if (bad == 0 || -(s + 1) > bad) bad = -(s + 1); }
        continue;
      }
      ro_process_samples(xz, yz, xy, sx, sy, sz, mask, 1, row);
      if (feats) memcpy(feats + s * F, row, sizeof(float) * (size_t)F);
      double dec[RO_MAX_CLASSES];
      ro_decision_one(row, m, kv, dec);
      double* p = proba + s * C;
      ro_platt(dec, m, p);
      int best = 0;
      for (int c = 1; c < C; ++c) if (p[c] > p[best]) best = c;      /* np.argmax: first maximum */
      label[s] = best;
      if (known) known[s] = p[best] >= min_proba ? 1 : 0;
    }
    free(xz);
    free(kv);
  }
  return bad;
}

/* common.py:106-121 calculate_matrix_indices (+ :93-97 cartesian_to_spherical): int() truncates. */
void ro_matrix_indices(double x, double y, double z, int sx, int sy, int sz, double r_min, double r_max,
                       double th_min, double th_max, double ph_min, double ph_max, int32_t* ijk) {
  const double r = sqrt(x * x + y * y + z * z);
  const double k180_pi = 180.0 / 3.14159265358979323846;
  const double phi = atan2(y, z) * k180_pi;
  const double theta = asin(x / r) * k180_pi;
  ijk[0] = (int32_t)((theta - th_min) * (sx - 1) / (th_max - th_min));
  ijk[1] = (int32_t)((phi - ph_min) * (sy - 1) / (ph_max - ph_min));
  ijk[2] = (int32_t)((r - r_min) * (sz - 1) / (r_max - r_min));
}
