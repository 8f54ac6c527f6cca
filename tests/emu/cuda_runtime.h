// stands in for <cuda_runtime.h> in the emulation build (everything is in cuda_emu.h, force-included)
#pragma once
