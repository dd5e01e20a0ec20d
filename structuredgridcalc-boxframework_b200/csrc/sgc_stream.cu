// sgc_stream.cu — placeholder until the streaming kernel lands (see next commit).
#include "sgc_device.h"
namespace sgc {
int launch_heat_stream(sgc_level_s*, sgc_level_s*, double, int, int, int, cudaStream_t) { return 1; }
}  // namespace sgc
