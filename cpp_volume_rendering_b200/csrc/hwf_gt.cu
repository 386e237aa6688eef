// hwf_gt.cu -- the rc1pcrtgt marcher in VRB_FILTER_HARDWARE mode: every volume tap (primary and secondary rays) is a
// texture-unit trilinear fetch, as in the reference's GL path.  Same code as march_gt.cu (march_gt_body.cuh), compiled
// WITH fp contraction: within the parity tolerance, not bit-exact.
#include "vrb_internal.cuh"
#include "march_gt_common.cuh"
// secondary rays marched together by one thread (measured on B200 at cfg4: 986 / 895 / 918 ms for 1 / 2 / 4 rays in flight)
#ifndef GT_ILP
#define GT_ILP 2
#endif
#define GT_HW 1
namespace gt_hw {
#include "march_gt_body.cuh"
}

int vrb_gt_launch_hw(vrb_ctx* c, const vrb_camera* cam, const GtConst& C, int count_samples) {
  return gt_hw::gt_launch(c, cam, C, count_samples);
}
