// cusim: placeholder for <cuda_fp16.h> — the sources built for the emulator do not use half precision.
#pragma once
struct __half { unsigned short x; };
struct __half2 { __half x, y; };
