// hwf_dos.cu -- the rc1pdosct marcher in VRB_FILTER_HARDWARE mode: volume and extinction pyramid are sampled by the
// texture units (trilinear, linear between mip levels), as the reference's GL samplers do.  Same code as march_dos.cu
// (march_dos_body.cuh) but compiled WITH fp contraction: this mode is within the parity tolerance, not bit-exact.
#include "vrb_internal.cuh"
#include "march_dos_common.cuh"
#define DOS_HW 1
namespace dos_hw {
#include "march_dos_body.cuh"
}

int vrb_dos_launch_hw(vrb_ctx* c, const vrb_camera* cam, const DosConst& C, int count_samples) {
  return dos_hw::dos_launch(c, cam, C, count_samples);
}

int vrb_dos_light_cache_launch_hw(vrb_ctx* c, const DosConst& C, const float eye_up[3], int rw, int rh, int rd) {
  return dos_hw::dos_light_cache_launch(c, C, eye_up, rw, rh, rd);
}
