//  Stand-in for <cuda_runtime.h> when a kernel source is compiled for the cusim host emulator.
#pragma once
#include "cusim.h"
