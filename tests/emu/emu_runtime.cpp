// TEST INFRASTRUCTURE ONLY: the one translation unit that defines the fiber switch and the block scheduler.
#define B2P_EMU_DEFINE_SWITCH
#include "cuda_emu.hpp"
