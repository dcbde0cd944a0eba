#include "bootstrap.h"

namespace vb {
namespace boot {

bool bootstrap_from_flow(const float*, int, int, const float*, float*, float*, float*) {
    printf("voldor_b200: monocular bootstrap is not available yet; provide depth priors or "
           "vb_set_bootstrap_override()\n");
    return false;
}

}  // namespace boot
}  // namespace vb
