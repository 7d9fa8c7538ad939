// Association costs of the tracker that consumes the tracking head's detections.
//
// Replaces, for one frame, the two cost matrices centernet_lightning/models/tracker.py:157-173 builds on the host:
//   reid cost  : scipy.spatial.distance.cdist(det_embeddings, track_embeddings, "cosine")            (tracker.py:61,157)
//   box cost   : 1 - IoU or 1 - GIoU of every (detection, track) box pair, utils/box.py:49-92         (tracker.py:63,169)
// The Hungarian assignment itself (scipy.optimize.linear_sum_assignment, tracker.py:28) stays on the host.
//
// Arithmetic is fp64 with the operation order of the host code (sequential dot products, no FMA contraction), so the
// matrices agree with scipy / numpy float64 to the last bit for float64 inputs.  One thread per (detection, track) pair;
// the matrices are tiny (k <= 1024 detections x T tracks), the point is to keep the per-frame path on the device.
#include <cuda_runtime.h>
#include <cstdint>
#include "cnl_common.h"

namespace cnl {

__device__ __forceinline__ double dot_seq(const double* a, const double* b, int n) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s = __dadd_rn(s, __dmul_rn(a[i], b[i]));
  return s;
}

// norms[i] = sqrt(sum x_i^2) for the rows of a (na) followed by the rows of b (nb)   (scipy precomputes them the same way)
__global__ void track_norms_kernel(const double* __restrict__ a, int na, const double* __restrict__ b, int nb, int dim,
                                   double* __restrict__ norms) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= na + nb) return;
  const double* x = (t < na) ? a + (size_t)t * dim : b + (size_t)(t - na) * dim;
  norms[t] = sqrt(dot_seq(x, x, dim));
}

__global__ void track_cost_kernel(const double* __restrict__ det_emb, const double* __restrict__ trk_emb, int dim,
                                  const double* __restrict__ norms, const double* __restrict__ det_box,
                                  const double* __restrict__ trk_box, int na, int nb, int giou,
                                  double* __restrict__ reid_cost, double* __restrict__ box_cost) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= na * nb) return;
  const int i = t / nb, j = t - i * nb;
  if (reid_cost != nullptr) {
    // scipy cdist "cosine": cosine = u.v / (|u| |v|), clipped to [-1, 1]; distance = 1 - cosine
    double c = dot_seq(det_emb + (size_t)i * dim, trk_emb + (size_t)j * dim, dim) / __dmul_rn(norms[i], norms[na + j]);
    if (fabs(c) > 1.0) c = copysign(1.0, c);
    reid_cost[t] = __dsub_rn(1.0, c);
  }
  if (box_cost != nullptr) {
    const double* p = det_box + 4 * i;
    const double* q = trk_box + 4 * j;
    // utils/box.py:49-61
    const double area1 = __dmul_rn(__dsub_rn(p[2], p[0]), __dsub_rn(p[3], p[1]));
    const double area2 = __dmul_rn(__dsub_rn(q[2], q[0]), __dsub_rn(q[3], q[1]));
    const double w = fmax(__dsub_rn(fmin(p[2], q[2]), fmax(p[0], q[0])), 0.0);
    const double h = fmax(__dsub_rn(fmin(p[3], q[3]), fmax(p[1], q[1])), 0.0);
    const double inter = __dmul_rn(w, h);
    const double uni = __dsub_rn(__dadd_rn(area1, area2), inter);
    const double iou = inter / uni;
    double v = iou;
    if (giou) {                                  // utils/box.py:70-80
      const double wi = fmax(__dsub_rn(fmax(p[2], q[2]), fmin(p[0], q[0])), 0.0);
      const double hi = fmax(__dsub_rn(fmax(p[3], q[3]), fmin(p[1], q[1])), 0.0);
      const double areai = __dmul_rn(wi, hi);
      v = __dsub_rn(iou, __dsub_rn(areai, uni) / areai);
    }
    box_cost[t] = __dsub_rn(1.0, v);
  }
}

}  // namespace cnl

using namespace cnl;

extern "C" size_t cnl_track_workspace_bytes(int n_det, int n_trk) {
  if (n_det < 0 || n_trk < 0) return 0;
  return align_up((size_t)(n_det + n_trk) * sizeof(double), 256);
}

extern "C" int cnl_track_cost_matrices(const double* det_emb, const double* trk_emb, int emb_dim,
                                       const double* det_box, const double* trk_box, int n_det, int n_trk, int giou,
                                       double* reid_cost, double* box_cost, void* workspace, size_t workspace_bytes, void* stream) {
  if (n_det <= 0 || n_trk <= 0) return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_track_cost_matrices: empty matrix (%d x %d)", n_det, n_trk);
  if ((reid_cost != nullptr) != (det_emb != nullptr && trk_emb != nullptr) || (reid_cost && emb_dim <= 0))
    return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_track_cost_matrices: embeddings, emb_dim and reid_cost must be given together");
  if ((box_cost != nullptr) != (det_box != nullptr && trk_box != nullptr))
    return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_track_cost_matrices: boxes and box_cost must be given together");
  if (!reid_cost && !box_cost) return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_track_cost_matrices: nothing to compute");
  if (reid_cost && (!workspace || workspace_bytes < cnl_track_workspace_bytes(n_det, n_trk)))
    return fail(CNL_ERR_WORKSPACE, "cnl_track_cost_matrices: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* norms = static_cast<double*>(workspace);
  if (reid_cost) track_norms_kernel<<<(n_det + n_trk + 127) / 128, 128, 0, st>>>(det_emb, n_det, trk_emb, n_trk, emb_dim, norms);
  const int total = n_det * n_trk;
  track_cost_kernel<<<(total + 127) / 128, 128, 0, st>>>(det_emb, trk_emb, emb_dim, norms, det_box, trk_box, n_det, n_trk, giou,
                                                        reid_cost, box_cost);
  CNL_CUDA_CHECK(cudaGetLastError());
  return CNL_OK;
}
