// hwf_vct.cu -- the rc1pvctsg marcher in VRB_FILTER_HARDWARE mode: volume, super-voxel pyramid (RG16F, trilinear +
// linear between levels) and pre-integration LUT (R16F, bilinear) are sampled by the texture units, as the reference's
// GL samplers do.  Same code as march_vct.cu (march_vct_body.cuh), compiled WITH fp contraction: within the parity
// tolerance, not bit-exact.
#include "vrb_internal.cuh"
#include "march_vct_common.cuh"
#define VCT_HW 1
namespace vct_hw {
#include "march_vct_body.cuh"
}

int vrb_vct_launch_hw(vrb_ctx* c, const vrb_camera* cam, const VctConst& C, int count_samples) {
  return vct_hw::vct_launch(c, cam, C, count_samples);
}

int vrb_vct_light_cache_launch_hw(vrb_ctx* c, const VctConst& C, int rw, int rh, int rd) {
  return vct_hw::vct_light_cache_launch(c, C, rw, rh, rd);
}

int vrb_vct_brick_launch_hw(vrb_ctx* c, const vrb_camera* cam, const VctConst& C, const VctBrick& B, const VctFront& front, int mode, int count_samples) {
  return vct_hw::vct_brick_launch(c, cam, C, B, front, mode, count_samples);
}
