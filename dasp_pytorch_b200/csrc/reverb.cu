// noise-shaped reverberation (placeholder: filled in by the reverb milestone)
#include "common.cuh"
namespace dasp {
void reverb_shutdown() {}
}  // namespace dasp
