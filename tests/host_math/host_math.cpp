// TEST-ONLY: compiles kaolin_b200/csrc/dibr_math.cuh as plain C++ (g++,
// -ffp-contract=off) and drives it with brute-force loops, so the exact
// arithmetic the CUDA kernels use can be compared with the oracle on a CPU-only
// box (tests/test_host_math.py).  Faces are visited in REVERSE order with the
// explicit (z, lowest-index) tie-break the tile kernels use, to prove the
// result does not depend on the visiting order.
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include "../../kaolin_b200/csrc/dibr_math.cuh"

using namespace dibr;

extern "C" {

void hm_bbox_to_rect(float multiplier, int W, int H, const float* bbox, int n, int* rects) {
  PixelGrid g = make_grid(multiplier, W, H);
  for (int i = 0; i < n; i++) {
    PixRect r = bbox_to_rect(g, bbox[4 * i], bbox[4 * i + 1], bbox[4 * i + 2], bbox[4 * i + 3]);
    rects[4 * i] = r.x_lo; rects[4 * i + 1] = r.x_hi; rects[4 * i + 2] = r.y_lo; rects[4 * i + 3] = r.y_hi;
  }
}

void hm_rasterize_forward(int B, int H, int W, int D, const float* z, const float* xy,
                          const float* bbox, const float* feat, const int64_t* first,
                          float multiplier, float eps, int64_t* sel, float* wout, float* out) {
  PixelGrid g = make_grid(multiplier, W, H);
  RasterConst rc = make_raster_const(eps);
  for (int b = 0; b < B; b++) {
    const int f0 = (int)first[b], f1 = (int)first[b + 1];
    PixRect* rects = (PixRect*)malloc(sizeof(PixRect) * (size_t)(f1 - f0 + 1));
    for (int f = f0; f < f1; f++)
      rects[f - f0] = bbox_to_rect(g, bbox[4 * f], bbox[4 * f + 1], bbox[4 * f + 2], bbox[4 * f + 3]);
    for (int iy = 0; iy < H; iy++) for (int ix = 0; ix < W; ix++) {
      const float x0 = pix_x(g, ix), y0 = pix_y(g, iy);
      float best_z = -INFINITY, bw0 = 0, bw1 = 0, bw2 = 0;
      int best_f = -1;
      for (int f = f1 - 1; f >= f0; f--) {
        const PixRect& r = rects[f - f0];
        if (ix < r.x_lo || ix >= r.x_hi || iy < r.y_lo || iy >= r.y_hi) continue;
        const float* p = xy + 6 * (int64_t)f;
        float w0, w1, w2;
        if (!raster_weights(rc, x0, y0, p[0], p[1], p[2], p[3], p[4], p[5], w0, w1, w2)) continue;
        const float zz = raster_depth(z[3 * f], z[3 * f + 1], z[3 * f + 2], w0, w1, w2);
        const bool take = !(zz <= best_z) || (zz == best_z && f < best_f);
        if (take) { best_z = zz; best_f = f; bw0 = w0; bw1 = w1; bw2 = w2; }
      }
      const int64_t pix = ((int64_t)b * H + iy) * W + ix;
      sel[pix] = best_f < 0 ? -1 : best_f - f0;
      wout[3 * pix] = bw0; wout[3 * pix + 1] = bw1; wout[3 * pix + 2] = bw2;
      for (int d = 0; d < D; d++) {
        if (best_f < 0) { out[pix * D + d] = 0.f; continue; }
        const float* ff = feat + (int64_t)best_f * 3 * D;
        out[pix * D + d] = raster_interp(ff[d], ff[D + d], ff[2 * D + d], bw0, bw1, bw2);
      }
    }
    free(rects);
  }
}

void hm_soft_mask_forward(int B, int H, int W, int F, int K, const float* xy, const float* bbox,
                          const int64_t* sel, float sigmainv, float multiplier, float* soft,
                          float* prob, int64_t* cidx, uint8_t* ctype) {
  PixelGrid g = make_grid(multiplier, W, H);
  PixRect* rects = (PixRect*)malloc(sizeof(PixRect) * (size_t)(F + 1));
  for (int b = 0; b < B; b++) {
    for (int f = 0; f < F; f++) {
      const float* bb = bbox + 4 * ((int64_t)b * F + f);
      rects[f] = bbox_to_rect(g, bb[0], bb[1], bb[2], bb[3]);
    }
    for (int iy = 0; iy < H; iy++) for (int ix = 0; ix < W; ix++) {
      const int64_t pix = ((int64_t)b * H + iy) * W + ix;
      for (int k = 0; k < K; k++) { prob[pix * K + k] = 0.f; cidx[pix * K + k] = -1; ctype[pix * K + k] = 0; }
      if (sel[pix] >= 0) { soft[pix] = 1.0f; continue; }
      const float x0 = pix_x(g, ix), y0 = pix_y(g, iy);
      int kid = 0;
      float allprob = 1.0f;
      for (int f = 0; f < F && kid < K; f++) {
        const PixRect& r = rects[f];
        if (ix < r.x_lo || ix >= r.x_hi || iy < r.y_lo || iy >= r.y_hi) continue;
        int edgeid;
        const float d2 = soft_min_dist(x0, y0, xy + 6 * ((int64_t)b * F + f), multiplier, edgeid);
        const float p = soft_prob(d2, sigmainv, multiplier);
        prob[pix * K + kid] = p; cidx[pix * K + kid] = f; ctype[pix * K + kid] = (uint8_t)(edgeid + 1);
        allprob = soft_accumulate(allprob, p);
        kid++;
      }
      soft[pix] = soft_finish(allprob);
    }
  }
  free(rects);
}

void hm_soft_mask_backward(int B, int H, int W, int F, int K, const float* gsoft, const float* soft,
                           const int64_t* sel, const float* prob, const int64_t* cidx,
                           const uint8_t* ctype, const float* xy, float sigmainv, float multiplier,
                           float* gxy) {
  PixelGrid g = make_grid(multiplier, W, H);
  const int64_t n = (int64_t)B * F * 6;
  double* acc = (double*)calloc((size_t)n, sizeof(double));
  for (int64_t pix = 0; pix < (int64_t)B * H * W; pix++) {
    if (sel[pix] >= 0) continue;
    const int ix = (int)(pix % W), iy = (int)((pix / W) % H), b = (int)(pix / ((int64_t)W * H));
    const float x0 = pix_x(g, ix), y0 = pix_y(g, iy);
    for (int k = 0; k < K; k++) {
      const int f = (int)cidx[pix * K + k];
      if (f < 0) break;
      float t[6];
      const int64_t base = ((int64_t)b * F + f) * 6;
      soft_backward_terms(x0, y0, xy + base, (int)ctype[pix * K + k] - 1, prob[pix * K + k],
                          soft[pix], gsoft[pix], sigmainv, multiplier, t);
      for (int j = 0; j < 6; j++) acc[base + j] += (double)t[j];
    }
  }
  for (int64_t i = 0; i < n; i++) gxy[i] = (float)acc[i];
  free(acc);
}

void hm_rasterize_backward(int B, int H, int W, int F, int D, const float* g, const int64_t* sel,
                           const float* w, const float* xy, const float* ff, float eps,
                           float* gxy, float* gff) {
  const int64_t nxy = (int64_t)B * F * 6, nff = (int64_t)B * F * 3 * D;
  double* axy = (double*)calloc((size_t)nxy, sizeof(double));
  double* aff = (double*)calloc((size_t)nff, sizeof(double));
  for (int64_t pix = 0; pix < (int64_t)B * H * W; pix++) {
    const int f = (int)sel[pix];
    if (f < 0) continue;
    const int64_t b = pix / ((int64_t)H * W);
    const int64_t face = b * F + f;
    const float* gp = g + pix * D;
    const float w0 = w[3 * pix], w1 = w[3 * pix + 1], w2 = w[3 * pix + 2];
    RasterBwdGeom G;
    raster_backward_geom(xy + face * 6, w0, w1, w2, eps, G);
    const float* c = ff + face * 3 * D;
    for (int d = 0; d < D; d++) {
      float t6[6];
      raster_backward_feature(G, gp[d], c[d], c[D + d], c[2 * D + d], t6);
      for (int j = 0; j < 6; j++) axy[face * 6 + j] += (double)t6[j];
      aff[face * 3 * D + d] += (double)(gp[d] * w0);
      aff[face * 3 * D + D + d] += (double)(gp[d] * w1);
      aff[face * 3 * D + 2 * D + d] += (double)(gp[d] * w2);
    }
  }
  for (int64_t i = 0; i < nxy; i++) gxy[i] = (float)axy[i];
  for (int64_t i = 0; i < nff; i++) gff[i] = (float)aff[i];
  free(axy); free(aff);
}

}  // extern "C"
