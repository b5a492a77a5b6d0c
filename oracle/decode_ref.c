/* TEST INFRASTRUCTURE -- plain-C restatement of the reference's heatmap decode, written as the
 * literal sequence of tensor ops the reference builds (double precision).  PARITY UNPINNED: the
 * reference holds no golden vectors; see oracle/metro_oracle.py.
 *
 *   volumetric.py:227-235  [N,H,W,D*J] (c = d*J + j) -> [N,J,H,W,D]; softmax over (H,W,D); decode [3,2,4]
 *   tfu.py:466-471         softmax = exp(x - max) / sum(exp(x - max))
 *   tfu.py:474-499         per axis: marginal over the other two axes, dot linspace(0,1,n)
 *   volumetric.py:288-306  xy = (c*lrc + stride/2) * box/proc_side ; z = c*box
 *   tfu3d.py:23-25         subtract the last joint
 *   main.py:127            gather(permutation)
 * Built by oracle/Makefile into oracle/_build/libdecode_ref.so; only tests/ load it. */
#include <math.h>
#include <stdlib.h>

static double linspace01(int i, int n) { return n > 1 ? (double)i / (double)(n - 1) : 0.0; }

int metro_oracle_decode(const float *head, int n, int side, int joints, int depth, int stride, int centered,
                        int proc_side, double box_mm, const int *perm, int n_out, double *out) {
  const int H = side, W = side, C = depth * joints;
  double *metric = (double *)malloc(sizeof(double) * joints * 3);
  double *p = (double *)malloc(sizeof(double) * H * W * depth);
  if (!metric || !p) return 1;
  const int last = proc_side - 1;
  const double lrc = (double)(last - (last % stride) - 1);
  for (int b = 0; b < n; ++b) {
    for (int j = 0; j < joints; ++j) {
      /* transposed[b, j, h, w, d] = head[b, h, w, d*J + j] */
      double mx = -INFINITY, sum = 0.0;
      for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w)
          for (int d = 0; d < depth; ++d) {
            const double v = head[(((size_t)b * H + h) * W + w) * C + d * joints + j];
            p[(h * W + w) * depth + d] = v;
            if (v > mx) mx = v;
          }
      for (int i = 0; i < H * W * depth; ++i) { p[i] = exp(p[i] - mx); sum += p[i]; }
      for (int i = 0; i < H * W * depth; ++i) p[i] /= sum;
      double cx = 0.0, cy = 0.0, cz = 0.0;
      for (int w = 0; w < W; ++w) {          /* axis 3 (W) -> x */
        double m = 0.0;
        for (int h = 0; h < H; ++h) for (int d = 0; d < depth; ++d) m += p[(h * W + w) * depth + d];
        cx += linspace01(w, W) * m;
      }
      for (int h = 0; h < H; ++h) {          /* axis 2 (H) -> y */
        double m = 0.0;
        for (int w = 0; w < W; ++w) for (int d = 0; d < depth; ++d) m += p[(h * W + w) * depth + d];
        cy += linspace01(h, H) * m;
      }
      for (int d = 0; d < depth; ++d) {      /* axis 4 (D) -> z */
        double m = 0.0;
        for (int h = 0; h < H; ++h) for (int w = 0; w < W; ++w) m += p[(h * W + w) * depth + d];
        cz += linspace01(d, depth) * m;
      }
      double ix = cx * lrc, iy = cy * lrc;
      if (centered) { ix += stride / 2; iy += stride / 2; }
      metric[3 * j] = ix * box_mm / proc_side;
      metric[3 * j + 1] = iy * box_mm / proc_side;
      metric[3 * j + 2] = cz * box_mm;
    }
    for (int o = 0; o < n_out; ++o)
      for (int a = 0; a < 3; ++a)
        out[((size_t)b * n_out + o) * 3 + a] = metric[3 * perm[o] + a] - metric[3 * (joints - 1) + a];
  }
  free(metric);
  free(p);
  return 0;
}
