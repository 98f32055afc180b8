// placeholder until the frame pipeline lands
#include "vo_internal.h"
namespace vo { void frame_plan_destroy(FramePlan*) {} }
extern "C" {
int vo_frames(vo_ctx*, const uint8_t*, const uint8_t*, int, int, int, const double*, const double*, const vo_frames_opts*, double*, int*, int*) {
  vo::set_error("vo_frames: not built yet"); return VO_ERR_STATE; }
}
