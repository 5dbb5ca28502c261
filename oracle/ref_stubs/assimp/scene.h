// Stub for oracle/ref_shim.cpp: the one assimp type StratusMath.h names (in ToMat4, which the shim never calls).
#pragma once
struct aiMatrix4x4 {
    float m[4][4];
    aiMatrix4x4 Transpose() const { return *this; }
    float* operator[](unsigned long i) { return m[i]; }
};
