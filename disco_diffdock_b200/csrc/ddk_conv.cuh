// Per-level constants of the tensor-product convolution (FasterTensorProduct basis, /root/reference/models/tensor_layers.py:39-116).
#pragma once

#include "ddk_device.cuh"

namespace ddk {

// LV = basis level = min(layer, 3): U basis rows in kernel order [0e | 1o comp-major | 1e comp-major | 0o] (constant
// factors are folded into the packed weights), DINP = input feature width rounded up to 4
template <int LV>
struct AccCfg {
  static constexpr int U = LV == 0 ? 96 : (LV == 1 ? 138 : (LV == 2 ? 180 : 276));
  static constexpr int DINP = LV == 0 ? 24 : (LV == 1 ? 44 : (LV == 2 ? 60 : 84));
};

}  // namespace ddk
