// Host build of gemmgan_b200/csrc/evalmetrics.cu (see emu.h): same C-ABI entry points, kernels run as fibers.
#include "emu.h"

#include "../../gemmgan_b200/csrc/evalmetrics.cu"
