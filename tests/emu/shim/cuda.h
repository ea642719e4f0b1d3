// <cuda.h> (driver API: tensor maps) is only used by the tcgen05 GEMM launcher, which the emulated build replaces.
#pragma once
