/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, fp32 with the reference's fp64 spots) of the four
 * CUDA kernels on Kaolin's DIB-R hot path.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this file's
 * library; the product (kaolin_b200) never does.
 *
 * Reference (NVIDIAGameWorks/kaolin v0.18.0) files restated here:
 *   kaolin/csrc/render/mesh/rasterization_cuda.cu:43-192   (rasterize forward)
 *   kaolin/csrc/render/mesh/rasterization_cuda.cu:238-402  (rasterize backward)
 *   kaolin/csrc/render/mesh/dibr_soft_mask_cuda.cu:27-184  (soft-mask forward)
 *   kaolin/csrc/render/mesh/dibr_soft_mask_cuda.cu:230-353 (soft-mask backward)
 *
 * Parity pinning: see oracle/README.md — checked against the reference's golden
 * fixtures tests/samples/dibr/{simple,sphere}/ *.pt (copied to tests/golden by
 * tests/golden/make_golden.py) and, on a GPU box, against oracle/_ref (the
 * reference's own .cu files compiled in place).
 *
 * The reference has no CPU implementation of this path (the C++ entry points
 * raise KAOLIN_NO_CUDA_ERROR, rasterization.cpp:95-102), so the arithmetic that
 * decides discrete outputs (face_idx, close_face_idx, dist_type) follows the
 * expression tree nvcc 12.9 -O3 generates for the <float> instantiations
 * (FMA contraction read from the SASS of oracle/_ref/ *.cu.o): products that
 * nvcc fuses are written with fmaf() and this file MUST be compiled with
 * -ffp-contract=off so gcc adds no fusion of its own.
 *
 * Backward passes: the reference accumulates with float atomicAdd in a
 * non-deterministic order; here every per-pixel term is computed in fp32 as
 * the reference does and the terms are accumulated in double (order-free to
 * ~1e-16), which is the natural "exact" target for a tolerance comparison.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SOFT_EPS 1e-7 /* dibr_soft_mask_cuda.cu:23 — a double literal */

/* ------------------------------------------------------------------------- */
/* rasterization_cuda.cu:85-86 — pixel centre, computed in C float even for  */
/* double tensors because `multiplier` is a `const float`.                   */
static inline float pix_x(float multiplier, int width, int ix) {
  return multiplier / (float)width * (float)(2 * ix + 1 - width);
}
static inline float pix_y(float multiplier, int height, int iy) {
  return multiplier / (float)height * (float)(height - 2 * iy - 1);
}

/*
 * rasterization_cuda.cu:43-192.  Inputs are the *packed* tensors the Python
 * wrapper builds (rasterization.py:308-327): only valid faces, coordinates
 * already multiplied, tight bboxes [xmin,ymin,xmax,ymax].
 * Outputs are fully written (the host wrapper's at::full(-1)/at::zeros,
 * rasterization.cpp:88-93, are folded in).
 */
void oracle_rasterize_forward_rows(
    int batch_size, int height, int width, int num_features,
    const float* face_vertices_z,      /* (NF,3)   */
    const float* face_vertices_image,  /* (NF,3,2) */
    const float* face_bboxes,          /* (NF,4)   */
    const float* face_features,        /* (NF,3,D) */
    const int64_t* first_idx_face_per_mesh, /* (B+1) */
    float multiplier, float eps,
    int64_t* selected_face_idx,        /* (B,H,W)   */
    float* output_weights,             /* (B,H,W,3) */
    float* interpolated_features,      /* (B,H,W,D) */
    int row0, int row1)                /* only image rows [row0,row1) are computed/written */
{
  const int D = num_features;
  for (int bidx = 0; bidx < batch_size; bidx++) {
    const int first_id_faces = (int)first_idx_face_per_mesh[bidx];
    const int last_id_faces = (int)first_idx_face_per_mesh[bidx + 1];
#pragma omp parallel for schedule(dynamic, 64)
    for (int pixel_idx = row0 * width; pixel_idx < row1 * width; pixel_idx++) {
      const int wididx = pixel_idx % width;
      const int heiidx = (pixel_idx - wididx) / width;
      float max_z0 = -INFINITY;
      int max_face_idx = -1;
      float max_w0 = 0.f, max_w1 = 0.f, max_w2 = 0.f;
      const int64_t totalidx1 = (int64_t)bidx * height * width + pixel_idx;
      const float x0 = pix_x(multiplier, width, wididx);
      const float y0 = pix_y(multiplier, height, heiidx);

      for (int face_idx = first_id_faces; face_idx < last_id_faces; face_idx++) {
        const float* bb = face_bboxes + (int64_t)face_idx * 4;
        /* :115 half-open bbox test */
        if (x0 < bb[0] || x0 >= bb[2] || y0 < bb[1] || y0 >= bb[3]) continue;
        const float* p = face_vertices_image + (int64_t)face_idx * 6;
        const float aex = p[0] - x0, aey = p[1] - y0;
        const float bex = p[2] - x0, bey = p[3] - y0;
        const float cex = p[4] - x0, cey = p[5] - y0;
        /* :131-133 with nvcc's contraction: fma(a,b,-(c*d)) */
        float w0 = fmaf(bex, cey, -(bey * cex));
        float w1 = fmaf(aey, cex, -(aex * cey));
        float w2 = fmaf(aex, bey, -(aey * bex));
        float norm = (w0 + w1) + w2;
        /* :140-142 eps is added in double, then rounded back to float */
        norm = (float)((double)norm + copysign((double)eps, (double)norm));
        w0 /= norm;
        w1 /= norm;
        w2 /= norm;
        /* :148 (NaN compares false => treated as inside) */
        if (w0 < 0.f || w1 < 0.f || w2 < 0.f) continue;
        const float* zz = face_vertices_z + (int64_t)face_idx * 3;
        /* :159 contraction: fma(cz,w2, fma(az,w0, bz*w1)) */
        const float z0 = fmaf(zz[2], w2, fmaf(zz[0], w0, zz[1] * w1));
        /* :162 strict: ties keep the first (lowest) face index */
        if (z0 <= max_z0) continue;
        max_z0 = z0;
        max_face_idx = face_idx;
        max_w0 = w0; max_w1 = w1; max_w2 = w2;
      }
      float* wout = output_weights + totalidx1 * 3;
      float* fout = interpolated_features + totalidx1 * D;
      if (max_face_idx > -1) {
        selected_face_idx[totalidx1] = max_face_idx - first_id_faces;
        wout[0] = max_w0; wout[1] = max_w1; wout[2] = max_w2;
        const float* ff = face_features + (int64_t)max_face_idx * 3 * D;
        for (int d = 0; d < D; d++) {
          /* :184-187 contraction: fma(r2,w2, fma(r0,w0, r1*w1)) */
          fout[d] = fmaf(ff[2 * D + d], max_w2, fmaf(ff[d], max_w0, ff[D + d] * max_w1));
        }
      } else {
        selected_face_idx[totalidx1] = -1;
        wout[0] = wout[1] = wout[2] = 0.f;
        for (int d = 0; d < D; d++) fout[d] = 0.f;
      }
    }
  }
}

void oracle_rasterize_forward(
    int batch_size, int height, int width, int num_features,
    const float* face_vertices_z, const float* face_vertices_image, const float* face_bboxes,
    const float* face_features, const int64_t* first_idx_face_per_mesh,
    float multiplier, float eps,
    int64_t* selected_face_idx, float* output_weights, float* interpolated_features)
{
  oracle_rasterize_forward_rows(batch_size, height, width, num_features, face_vertices_z,
                                face_vertices_image, face_bboxes, face_features,
                                first_idx_face_per_mesh, multiplier, eps, selected_face_idx,
                                output_weights, interpolated_features, 0, height);
}

/*
 * rasterization_cuda.cu:238-402.  face_vertices_image is UNSCALED here
 * (rasterization.py:347-350 saves the original tensor).  Accumulation in
 * double (see header).  Outputs (B,F,3,2) and (B,F,3,D), fully written.
 */
void oracle_rasterize_backward_rows(
    int batch_size, int height, int width, int num_faces, int feat_dim,
    const float* grad_interpolated_features, /* (B,H,W,D) */
    const int64_t* selected_face_idx,        /* (B,H,W), original face ids */
    const float* output_weights,             /* (B,H,W,3) */
    const float* face_vertices_image,        /* (B,F,3,2) */
    const float* face_features,              /* (B,F,3,D) */
    float eps,
    float* grad_face_vertices_image,         /* (B,F,3,2) */
    float* grad_face_features,               /* (B,F,3,D) */
    int row0, int row1)                      /* only pixels of rows [row0,row1) contribute */
{
  const int D = feat_dim;
  const int64_t n_xy = (int64_t)batch_size * num_faces * 6;
  const int64_t n_ff = (int64_t)batch_size * num_faces * 3 * D;
  double* acc_xy = (double*)calloc((size_t)n_xy, sizeof(double));
  double* acc_ff = (double*)calloc((size_t)n_ff, sizeof(double));
  const int64_t num_pixels = (int64_t)height * width;
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t true_pixel_idx = 0; true_pixel_idx < batch_size * num_pixels; true_pixel_idx++) {
    const int64_t batch_idx = true_pixel_idx / num_pixels;
    const int row = (int)((true_pixel_idx % num_pixels) / width);
    if (row < row0 || row >= row1) continue;
    const int face_idx = (int)selected_face_idx[true_pixel_idx];
    if (face_idx < 0) continue;
    const float* g = grad_interpolated_features + true_pixel_idx * D;
    const float* wv = output_weights + true_pixel_idx * 3;
    const int64_t true_face_idx = batch_idx * num_faces + face_idx;
    const int64_t start_image_idx = true_face_idx * 6;
    const int64_t start_features_idx = true_face_idx * 3 * D;
    /* :272-285 */
    for (int ii = 0; ii < 3; ii++) {
      for (int d = 0; d < D; d++) {
        const float val = g[d] * wv[ii];
#pragma omp atomic
        acc_ff[start_features_idx + ii * D + d] += (double)val;
      }
    }
    /* :292-374 — expression tree as compiled by nvcc 12.9 for the <float> kernel
     * (SASS offsets 0x20b0-0x27d0 of oracle/_ref/rasterization_cuda.cu.o).  The
     * partials are differences of nearly equal products divided by k3^2, so a
     * different (mathematically equal) association differs by ~1e-4 relative. */
    const float* p = face_vertices_image + start_image_idx;
    const float ax = p[0], ay = p[1], bx = p[2], by = p[3], cx = p[4], cy = p[5];
    const float aw = wv[0], bw = wv[1], cw = wv[2];
    const float pp = by - ay, n = cx - ax, m = bx - ax, q = cy - ay;
    float k3 = fmaf(m, q, -(pp * n));
    k3 = (float)((double)k3 + copysign((double)eps, (double)k3));
    const float y0 = fmaf(cy, cw, fmaf(ay, aw, by * bw));
    const float x0 = fmaf(cx, cw, fmaf(ax, aw, bx * bw));
    const float t = y0 - ay, s = x0 - ax;
    const float k1 = fmaf(q, s, -(n * t));
    const float k2 = fmaf(m, t, -(pp * s));
    const float z1 = 0.f * k1, z3 = 0.f * k3, z2 = 0.f * k2;
    const float tk3 = t * k3, sk3 = s * k3;
    const float dw1ds = fmaf(q, k3, -z1);     /* dk1ds*k3 - dk3ds*k1 */
    const float dw1dm = fmaf(-q, k1, z3);     /* dk1dm*k3 - dk3dm*k1 */
    const float dw2dm = fmaf(-q, k2, tk3);
    const float dw1dn = fmaf(pp, k1, -tk3);
    const float dw1dp = fmaf(n, k1, z3);
    const float dw1dq = fmaf(-m, k1, sk3);
    const float dw1dt = fmaf(-n, k3, -z1);
    const float dw2dp = fmaf(n, k2, -sk3);
    const float dw2ds = fmaf(-pp, k3, -z2);
    const float dw2dt = fmaf(m, k3, -z2);
    const float dw2dn = fmaf(pp, k2, z3);
    const float dw2dq = fmaf(-m, k2, z3);
    /* :362-374; the a-vertex partials are kept negated, as the compiled code does */
    const float n1ay = dw1dt + (dw1dp + dw1dq);
    const float n1ax = dw1ds + (dw1dm + dw1dn);
    const float n2ax = dw2ds + (dw2dm + dw2dn);
    const float n2ay = dw2dt + (dw2dp + dw2dq);
    const float k3sq = k3 * k3;

    const float* ff = face_features + start_features_idx;
    /* :376-399 */
    for (int d = 0; d < D; d++) {
      const float c0 = ff[d], c1 = ff[D + d], c2 = ff[2 * D + d];
      const float d1 = c1 - c0, d2 = c2 - c0;
      const float dldI = g[d] / k3sq;
      const float v[6] = {
          fmaf(-n2ax, d2, -(n1ax * d1)) * dldI, fmaf(-n2ay, d2, -(n1ay * d1)) * dldI,
          fmaf(dw1dm, d1, dw2dm * d2) * dldI,   fmaf(dw1dp, d1, dw2dp * d2) * dldI,
          fmaf(dw1dn, d1, dw2dn * d2) * dldI,   fmaf(dw1dq, d1, dw2dq * d2) * dldI};
      for (int j = 0; j < 6; j++) {
#pragma omp atomic
        acc_xy[start_image_idx + j] += (double)v[j];
      }
    }
  }
  for (int64_t i = 0; i < n_xy; i++) grad_face_vertices_image[i] = (float)acc_xy[i];
  for (int64_t i = 0; i < n_ff; i++) grad_face_features[i] = (float)acc_ff[i];
  free(acc_xy);
  free(acc_ff);
}

void oracle_rasterize_backward(
    int batch_size, int height, int width, int num_faces, int feat_dim,
    const float* grad_interpolated_features, const int64_t* selected_face_idx,
    const float* output_weights, const float* face_vertices_image, const float* face_features,
    float eps, float* grad_face_vertices_image, float* grad_face_features)
{
  oracle_rasterize_backward_rows(batch_size, height, width, num_faces, feat_dim,
                                 grad_interpolated_features, selected_face_idx, output_weights,
                                 face_vertices_image, face_features, eps, grad_face_vertices_image,
                                 grad_face_features, 0, height);
}

/*
 * dibr_soft_mask_cuda.cu:27-184.  face_vertices_image is already multiplied
 * (dibr.py:32), face_large_bboxes = [min - boxlen*m, max + boxlen*m]
 * (dibr.py:33-39).  ALL faces take part (no validity mask).  K-list outputs are
 * fully written (padding -1 / 0 / 0 as dibr_soft_mask.cpp:86-97 allocates).
 * Any of close_face_prob / close_face_idx / close_face_dist_type may be NULL.
 */
void oracle_soft_mask_forward_rows(
    int batch_size, int height, int width, int num_faces, int knum,
    const float* face_vertices_image,  /* (B,F,3,2) * multiplier */
    const float* face_bboxes,          /* (B,F,4) enlarged */
    const int64_t* selected_face_idx,  /* (B,H,W) */
    float sigmainv, float multiplier,
    float* soft_mask,                  /* (B,H,W)   */
    float* close_face_prob,            /* (B,H,W,K) or NULL */
    int64_t* close_face_idx,           /* (B,H,W,K) or NULL */
    uint8_t* close_face_dist_type,     /* (B,H,W,K) or NULL */
    int row0, int row1)                /* only image rows [row0,row1) are computed/written */
{
  const int64_t P = (int64_t)batch_size * height * width;
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t totalidx1 = 0; totalidx1 < P; totalidx1++) {
    const int wididx = (int)(totalidx1 % width);
    const int heiidx = (int)((totalidx1 / width) % height);
    if (heiidx < row0 || heiidx >= row1) continue;
    const int bidx = (int)(totalidx1 / ((int64_t)width * height));
    const int64_t totalidxk = totalidx1 * knum;
    float probs_local[knum > 0 ? knum : 1];
    if (close_face_prob) for (int k = 0; k < knum; k++) close_face_prob[totalidxk + k] = 0.f;
    if (close_face_idx) for (int k = 0; k < knum; k++) close_face_idx[totalidxk + k] = -1;
    if (close_face_dist_type) for (int k = 0; k < knum; k++) close_face_dist_type[totalidxk + k] = 0;

    if (selected_face_idx[totalidx1] >= 0) { /* :68-70 */
      soft_mask[totalidx1] = 1.0f;
      continue;
    }
    const float x0 = pix_x(multiplier, width, wididx);
    const float y0 = pix_y(multiplier, height, heiidx);
    int kid = 0;
    for (int f = 0; f < num_faces && kid < knum; f++) { /* :80, :170-171 */
      const int64_t shift1 = (int64_t)bidx * num_faces + f;
      const float* bb = face_bboxes + shift1 * 4;
      if (x0 < bb[0] || x0 >= bb[2] || y0 < bb[1] || y0 >= bb[3]) continue; /* :95 */
      const float* v = face_vertices_image + shift1 * 6;
      float pdis[6];
      for (int i = 0; i < 3; i++) { /* :102-140 */
        const float x1 = v[i * 2], y1 = v[i * 2 + 1];
        const float x2 = v[((i + 1) % 3) * 2], y2 = v[((i + 1) % 3) * 2 + 1];
        const float A = y2 - y1;
        const float Bc = x1 - x2;
        const float C = fmaf(y1, x2, -(x1 * y2));          /* x2*y1 - x1*y2 */
        const float AA = A * A, BB = Bc * Bc, AB = A * Bc;
        const float down = AA + BB;
        const double down64 = (double)down + SOFT_EPS;
        const float up = C + fmaf(y0, Bc, x0 * A);         /* A*x0 + B*y0 + C */
        const float x3n = fmaf(-A, C, fmaf(x0, BB, -(y0 * AB)));
        const float y3n = fmaf(-Bc, C, fmaf(y0, AA, -(x0 * AB)));
        const float x3 = (float)((double)x3n / down64);
        const float y3 = (float)((double)y3n / down64);
        const float direct = fmaf(x3 - x1, x3 - x2, (y3 - y1) * (y3 - y2));
        if (direct > 0.f) {
          pdis[i] = (4.f * multiplier) * multiplier;
        } else {
          pdis[i] = (float)((double)(up * up) / down64);
        }
      }
      for (int i = 0; i < 3; i++) { /* :144-149 (nvcc: fma(dx,dx, dy*dy)) */
        const float dx = x0 - v[i * 2], dy = y0 - v[i * 2 + 1];
        pdis[i + 3] = fmaf(dx, dx, dy * dy);
      }
      int edgeid = 0;
      float dissquare = pdis[0];
      for (int i = 1; i < 6; i++) { /* :151-159 first minimum */
        if (dissquare > pdis[i]) { dissquare = pdis[i]; edgeid = i; }
      }
      const float z = sigmainv * dissquare / multiplier / multiplier;
      const float prob = expf(-z);
      probs_local[kid] = prob;
      if (close_face_prob) close_face_prob[totalidxk + kid] = prob;
      if (close_face_idx) close_face_idx[totalidxk + kid] = f;
      if (close_face_dist_type) close_face_dist_type[totalidxk + kid] = (uint8_t)(edgeid + 1);
      kid++;
    }
    float allprob = 1.0f; /* :174-182: double product rounded to float each step */
    for (int i = 0; i < kid; i++) allprob = (float)((double)allprob * (1.0 - (double)probs_local[i]));
    soft_mask[totalidx1] = (float)(1.0 - (double)allprob);
  }
}

void oracle_soft_mask_forward(
    int batch_size, int height, int width, int num_faces, int knum,
    const float* face_vertices_image, const float* face_bboxes, const int64_t* selected_face_idx,
    float sigmainv, float multiplier, float* soft_mask, float* close_face_prob,
    int64_t* close_face_idx, uint8_t* close_face_dist_type)
{
  oracle_soft_mask_forward_rows(batch_size, height, width, num_faces, knum, face_vertices_image,
                                face_bboxes, selected_face_idx, sigmainv, multiplier, soft_mask,
                                close_face_prob, close_face_idx, close_face_dist_type, 0, height);
}

/*
 * dibr_soft_mask_cuda.cu:230-353.  Per-term arithmetic as the reference
 * (float variables, double where the literals promote); accumulation in double.
 */
void oracle_soft_mask_backward_rows(
    int batch_size, int height, int width, int num_faces, int knum,
    const float* grad_soft_mask,        /* (B,H,W) */
    const float* soft_mask,             /* (B,H,W) */
    const int64_t* selected_face_idx,   /* (B,H,W) */
    const float* close_face_prob,       /* (B,H,W,K) */
    const int64_t* close_face_idx,      /* (B,H,W,K) */
    const uint8_t* close_face_dist_type,/* (B,H,W,K) */
    const float* face_vertices_image,   /* (B,F,3,2) * multiplier */
    float sigmainv, float multiplier,
    float* grad_face_vertices_image,    /* (B,F,3,2) */
    int row0, int row1)                 /* only pixels of rows [row0,row1) contribute */
{
  const int64_t n_xy = (int64_t)batch_size * num_faces * 6;
  double* acc = (double*)calloc((size_t)n_xy, sizeof(double));
  const int64_t P = (int64_t)batch_size * height * width;
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t totalidx1 = 0; totalidx1 < P; totalidx1++) {
    const int wididx = (int)(totalidx1 % width);
    const int heiidx = (int)((totalidx1 / width) % height);
    const int bidx = (int)(totalidx1 / ((int64_t)width * height));
    const int64_t totalidxk = totalidx1 * knum;
    if (heiidx < row0 || heiidx >= row1) continue;
    if (selected_face_idx[totalidx1] >= 0) continue;
    const float x0 = pix_x(multiplier, width, wididx);
    const float y0 = pix_y(multiplier, height, heiidx);
    const float dLdp = grad_soft_mask[totalidx1];
    const float allprob = soft_mask[totalidx1];
    for (int kid = 0; kid < knum; kid++) {
      const int f = (int)close_face_idx[totalidxk + kid];
      if (f < 0) break;
      const int64_t shift6 = ((int64_t)bidx * num_faces + f) * 6;
      const float prob = close_face_prob[totalidxk + kid];
      const float dLdz = (float)(-1.0 * sigmainv * dLdp * (1.0 - allprob)
                                 / (1.0 - prob + SOFT_EPS) * prob); /* :283-284 */
      const int edgeid = (int)close_face_dist_type[totalidxk + kid] - 1;
      if (edgeid >= 3) { /* :289-302 */
        const int64_t pshift = shift6 + (edgeid - 3) * 2;
        const float x1 = face_vertices_image[pshift], y1 = face_vertices_image[pshift + 1];
        const float dLdx1 = dLdz * 2 * (x1 - x0);
        const float dLdy1 = dLdz * 2 * (y1 - y0);
#pragma omp atomic
        acc[pshift] += (double)(dLdx1 / multiplier);
#pragma omp atomic
        acc[pshift + 1] += (double)(dLdy1 / multiplier);
      } else { /* :304-347, with the FMA contraction of the compiled <float> kernel */
        const int64_t pshift = shift6 + edgeid * 2;
        const int64_t pshift2 = shift6 + ((edgeid + 1) % 3) * 2;
        const float x1 = face_vertices_image[pshift], y1 = face_vertices_image[pshift + 1];
        const float x2 = face_vertices_image[pshift2], y2 = face_vertices_image[pshift2 + 1];
        const float A = y2 - y1, Bc = x1 - x2;
        const float C = fmaf(y1, x2, -(x1 * y2));
        const float up = C + fmaf(y0, Bc, x0 * A);
        const float down = fmaf(Bc, Bc, A * A);
        const double down64 = (double)down + SOFT_EPS;
        const float dissquare = (float)((double)(up * up) / down64);
        const float nA = fmaf(x0, up, -(A * dissquare));
        const float nB = fmaf(y0, up, -(Bc * dissquare));
        const float dzdA = (float)((double)(nA + nA) / down64);
        const float dzdB = (float)((double)(nB + nB) / down64);
        const float dzdC = (float)((double)(up + up) / down64);
        const float dLdx1 = dLdz * fmaf(-y2, dzdC, dzdB);
        const float dLdy1 = dLdz * fmaf(x2, dzdC, -dzdA);
        const float dLdx2 = dLdz * fmaf(y1, dzdC, -dzdB);
        const float dLdy2 = dLdz * fmaf(-x1, dzdC, dzdA);
#pragma omp atomic
        acc[pshift] += (double)(dLdx1 / multiplier);
#pragma omp atomic
        acc[pshift + 1] += (double)(dLdy1 / multiplier);
#pragma omp atomic
        acc[pshift2] += (double)(dLdx2 / multiplier);
#pragma omp atomic
        acc[pshift2 + 1] += (double)(dLdy2 / multiplier);
      }
    }
  }
  for (int64_t i = 0; i < n_xy; i++) grad_face_vertices_image[i] = (float)acc[i];
  free(acc);
}

void oracle_soft_mask_backward(
    int batch_size, int height, int width, int num_faces, int knum,
    const float* grad_soft_mask, const float* soft_mask, const int64_t* selected_face_idx,
    const float* close_face_prob, const int64_t* close_face_idx,
    const uint8_t* close_face_dist_type, const float* face_vertices_image,
    float sigmainv, float multiplier, float* grad_face_vertices_image)
{
  oracle_soft_mask_backward_rows(batch_size, height, width, num_faces, knum, grad_soft_mask,
                                 soft_mask, selected_face_idx, close_face_prob, close_face_idx,
                                 close_face_dist_type, face_vertices_image, sigmainv, multiplier,
                                 grad_face_vertices_image, 0, height);
}

/* Thread control for the cpu_baseline leg of bench.py. */
#ifdef _OPENMP
#include <omp.h>
int oracle_max_threads(void) { return omp_get_max_threads(); }
void oracle_set_threads(int n) { omp_set_num_threads(n); }
#else
int oracle_max_threads(void) { return 1; }
void oracle_set_threads(int n) { (void)n; }
#endif
