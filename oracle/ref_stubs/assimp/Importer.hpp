// Stub for oracle/ref_shim.cpp: StratusMath.h includes assimp; nothing the shim compiles uses it.
#pragma once
