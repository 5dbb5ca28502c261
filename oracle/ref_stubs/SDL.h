// Stub for oracle/ref_shim.cpp: StratusCommon.h includes SDL.h; nothing the shim compiles uses it.
#pragma once
