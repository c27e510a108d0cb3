// picked up instead of the CUDA toolkit header when the library is built for host-side emulation
#pragma once
#include "cuda_emu.h"
