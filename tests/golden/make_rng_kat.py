"""Generates tests/golden/rng_kat.json: known answers for the gather's RNG.

Independent restatement in pure Python integers of Source/Shaders/common.glsl:183-214
(hash(uint), hash(uvec3), floatConstruct, random(vec3)) — it shares no code with
oracle/ or the CUDA kernels.  Run once; the JSON is committed.

    python tests/golden/make_rng_kat.py
"""
import json
import os
import struct

M = 0xFFFFFFFF


def hash1(x):
    x = (x + ((x << 10) & M)) & M
    x ^= x >> 6
    x = (x + ((x << 3) & M)) & M
    x ^= x >> 11
    x = (x + ((x << 15) & M)) & M
    return x


def fbits(f):
    return struct.unpack("<I", struct.pack("<f", f))[0]


def hash3(x, y, z):
    return hash1(fbits(x) ^ hash1(fbits(y)) ^ hash1(fbits(z)))


def float_construct(m):
    m = (m & 0x007FFFFF) | 0x3F800000
    return struct.unpack("<f", struct.pack("<I", m))[0] - 1.0     # exact in binary64, result representable in binary32


cases = []
for (px, py) in [(0, 0), (1, 0), (0, 1), (319, 179), (639, 359), (1919, 1079), (3839, 2159), (7679, 4319)]:
    for t in (1000.0, 1016.0 + 10000.0):
        x, y = px + 0.5, py + 0.5
        h = hash3(x, y, t)
        cases.append({"x": x, "y": y, "time": t, "hash": h, "random_bits": fbits(float_construct(h))})
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rng_kat.json")
json.dump(cases, open(out, "w"), indent=1)
print(len(cases), "cases ->", out)
