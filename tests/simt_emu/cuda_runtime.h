// stub: the emulated build gets its CUDA vocabulary from simt_emu.h
#pragma once
#include "simt_emu.h"
