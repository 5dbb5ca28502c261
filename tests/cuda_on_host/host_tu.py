"""CUDA on the host — TEST INFRASTRUCTURE, never part of the product.

A kernel of stratusgfx_b200/csrc that reaches no barrier and no warp collective on the inputs of a test can be executed thread
by thread by ordinary C++: this module makes a host translation unit of it WITHOUT touching the sources.  The files of csrc/ are
copied into a build directory under tests/cuda_on_host/_build/ (git-ignored) with a handful of lexical rules applied, and g++
compiles a small wrapper that includes the copy:

  1. `kernel<<<grid, block, smem, stream>>>(args);`         ->  `SGI_HOST_LAUNCH(kernel, grid, block, smem, args);`  (cuda_host_shim.h:
     one call of the kernel function per thread, CTAs spread over the host's cores by OpenMP; or, with -DSGI_HOST_LOCKSTEP, one
     OS thread per CUDA thread of a CTA with working barriers, shuffles and votes)
  2. the one-line helpers whose body is inline PTX (`lg2.approx`, `ex2.approx`, `prefetch`, `fma.rn.f32x2`, `ld.acquire` ...)
     ->  dropped; cuda_host_shim.h defines functions of the same names with the IEEE / libm meaning
  3. `cudaFuncGetAttributes(&a, kernel) == cudaSuccess`      ->  `true`  (the preload helpers have nothing to load)
  4. `#ifdef __CUDACC__` in sgi_internal.h                   ->  `#if 1`  (the band helpers are compiled; with no neighbour attached
     they return before their barrier — in the thread-by-thread mode a barrier that IS reached aborts the process, see the shim)
  5. `extern __shared__ T name[];`                            ->  `T* name = (T*)sgihost::dynSmem();`
  6. sgi_band.cu                                             ->  its `namespace sgi` half only (the extern "C" half is CUDA IPC plumbing)
  7. k_atrous_tma.cu (whole-library build only)              ->  the mbarrier / cp.async.bulk.tensor helpers are dropped like the helpers of
     rule 2 and the two bare `asm volatile` statements of its kernel (`fence.mbarrier_init`, `prefetch.global.L2`) become empty
     statements; cuda_host_shim.h supplies a tensor map (its own layout inside the 128 opaque bytes of a CUtensorMap, encoded by a
     stand-in for cuTensorMapEncodeTiled that rejects what the driver rejects: a base or a stride that is not a multiple of 16 bytes,
     a box row that is not a multiple of 16 bytes, a box dimension above 256) and a tmaLoad4d that copies the box synchronously,
     zero-filling coordinates outside the tensor.  `#ifdef __CUDACC__` around the TMA path in k_denoise.cu and sgi_abi.cu becomes
     `#ifdef SGI_HOST_TMA`, which only the whole-library build defines: the per-kernel host tests keep atrousKernel on everything,
     the ABI-level host tests run the lattice tiles + border CTAs exactly as the GPU does.

Everything else — the kernel body, its index arithmetic, its exact-arithmetic chains — is compiled as written, with
-ffp-contract=off so that `a * b + c` is never fused behind the code's back.
"""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "stratusgfx_b200", "csrc")
BUILD = os.path.join(HERE, "_build")
HEADERS = ("sgi_internal.h", "sgi_math.cuh", "sgi_merge.cuh", "sgi_denoise.cuh")

# rule 2: helpers that the shim supplies (name -> only checked for presence, the definitions live in cuda_host_shim.h)
PTX_HELPERS = ("prefetchL2", "lg2Fast", "ex2Fast", "rcpFast", "rsqrtFast", "pack2", "unpack2", "fma2", "ldAcquireSys", "stReleaseSys", "sqrtApprox", "sqrtFast",
               "mbarInit", "mbarExpectTx", "mbarWait", "tmaLoad4d")
# rule 7: bare inline-PTX statements of k_atrous_tma.cu's kernel that have no host meaning
_BARE_ASM = re.compile(r'asm volatile\("(fence\.mbarrier_init\.release\.cluster|prefetch\.global\.L2 \[%0\]);"\s*:[^\n]*?\);')

_LAUNCH = re.compile(r"([A-Za-z_]\w*(?:<[^<>;]*>)?)\s*<<<((?:(?!<<<|>>>).)*)>>>\s*\(", re.S)


def _matching(text, start, open_ch, close_ch):
    """index just past the bracket that closes text[start] (string literals are skipped)"""
    depth, i, n = 0, start, len(text)
    while i < n:
        c = text[i]
        if c == '"':
            i += 1
            while text[i] != '"':
                i += 2 if text[i] == "\\" else 1
        elif c == open_ch:
            depth += 1
        elif c == close_ch:
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1
    raise ValueError("unbalanced " + open_ch)


def _split_top_level(s):
    parts, depth, cur = [], 0, ""
    for c in s:
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        if c == "," and depth == 0:
            parts.append(cur.strip()); cur = ""
        else:
            cur += c
    parts.append(cur.strip())
    return parts


def _drop_ptx_helpers(text):
    dropped = []
    for name in PTX_HELPERS:
        m = re.search(r"^[ \t]*(?:SGI_DEV|__device__ __forceinline__)[^\n;{]*\b" + name + r"\s*\(", text, re.M)
        if not m:
            continue
        brace = text.index("{", m.end())
        end = _matching(text, brace, "{", "}")
        body = text[brace:end]
        if "asm" not in body:
            raise ValueError(f"{name}: expected an inline-PTX body")
        text = text[:m.start()] + f"// [host emulation] {name}: defined by cuda_host_shim.h\n" + text[end:].lstrip("\n")
        dropped.append(name)
    return text, dropped


def _rewrite_launches(text):
    out, pos, count = "", 0, 0
    while True:
        m = _LAUNCH.search(text, pos)
        if not m:
            return out + text[pos:], count
        if "//" in text[text.rfind("\n", 0, m.start()) + 1:m.start()]:      # inside a comment
            out += text[pos:m.end()]; pos = m.end()
            continue
        cfg = _split_top_level(m.group(2))
        if len(cfg) < 2:
            raise ValueError("launch configuration: " + m.group(0))
        close = _matching(text, m.end() - 1, "(", ")")
        args = text[m.end():close - 1]
        smem = cfg[2] if len(cfg) > 2 else "0"
        out += text[pos:m.start()] + f"SGI_HOST_LAUNCH({m.group(1)}, {cfg[0]}, {cfg[1]}, {smem}, {args})"
        pos = close
        count += 1


def translate(text, name, whole_band_file=False):
    text, dropped = _drop_ptx_helpers(text)
    text, launches = _rewrite_launches(text)
    text = re.sub(r"cudaFuncGetAttributes\(&a,\s*[^)]*\)\s*==\s*cudaSuccess", "true", text)
    # rule 5: dynamic shared memory is a per-CTA allocation of the lockstep runtime
    text = re.sub(r"extern\s+__shared__\s+([\w:]+)\s+(\w+)\s*\[\s*\]\s*;", r"\1* \2 = (\1*)sgihost::dynSmem();", text)
    if name == "sgi_band.cu" and not whole_band_file:
        # rule 6: the extern "C" half of sgi_band.cu is CUDA IPC plumbing; the host run keeps the namespace sgi half (bandLinks,
        # armInKernelSignal, signalNeighbours, the flag kernels)
        text = text[:text.index("using namespace sgi;")]
    if name == "sgi_internal.h":
        text = text.replace("#ifdef __CUDACC__", "#if 1   // [host emulation] was: #ifdef __CUDACC__")
    if name in ("k_denoise.cu", "sgi_abi.cu"):
        text = text.replace("#ifdef __CUDACC__", "#ifdef SGI_HOST_TMA   // [host emulation] was: #ifdef __CUDACC__")
    if name == "k_atrous_tma.cu":
        text = _BARE_ASM.sub(lambda m: "/* [host emulation] " + m.group(1) + " */;", text)
    code = re.sub(r"//[^\n]*", "", text)
    if re.search(r"\basm\b", code):
        raise ValueError(f"{name}: inline PTX outside the known helpers — extend PTX_HELPERS and the shim")
    return text, dropped, launches


def make_build_dir(kernel_files, build=BUILD, whole_band_file=False):
    """Translated copies of the kernel files and the csrc headers -> tests/cuda_on_host/_build/; returns the directory."""
    if isinstance(kernel_files, str):
        kernel_files = (kernel_files,)
    os.makedirs(build, exist_ok=True)
    for f in HEADERS + tuple(kernel_files):
        src = open(os.path.join(CSRC, f)).read()
        text, _, _ = translate(src, f, whole_band_file)
        dst = os.path.join(build, f)
        if not os.path.exists(dst) or open(dst).read() != text:
            open(dst, "w").write(text)
    return build


def _gxx(sources, so, build, defines, extra=()):
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["g++", "-O2", *extra, "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-pthread", "-ffp-contract=off", "-w",
           "-I", build, "-I", HERE, "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include")] + \
          ["-D" + d for d in defines] + list(sources) + ["-o", so]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n" + r.stderr[-6000:])
    return so


def compile_wrapper(wrapper_cpps, kernel_files, out_name, defines=(), extra=()):
    """g++ the wrapper(s) (each includes one "_build/<kernel file>") into one shared object; rebuilt when any input is newer."""
    if isinstance(wrapper_cpps, str):
        wrapper_cpps = (wrapper_cpps,)
    if isinstance(kernel_files, str):
        kernel_files = (kernel_files,)
    build = make_build_dir(kernel_files)
    so = os.path.join(build, out_name)
    srcs = [os.path.join(HERE, w) for w in wrapper_cpps] + [os.path.join(HERE, "cuda_host_shim.h"), os.path.abspath(__file__)] + \
           [os.path.join(CSRC, f) for f in HEADERS + tuple(kernel_files)]
    if os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(s) for s in srcs):
        return so
    return _gxx([os.path.join(HERE, w) for w in wrapper_cpps], so, build, defines, extra)


ABI_FILES = ("sgi_abi.cu", "sgi_band.cu", "k_prepass.cu", "k_direct.cu", "k_ssao.cu", "k_taa.cu", "k_cubemaps.cu", "k_vpl.cu", "k_tiles.cu", "k_gather.cu", "k_denoise.cu",
             "k_atrous_tma.cu")


def compile_whole_library(sanitize=False):
    """EVERY translation unit of libstratusgi.so — the C ABI included — for the host, in lockstep mode and over the CUDA-runtime
    stand-ins of cuda_host_shim.h (-DSGI_HOST_RUNTIME): tests/cuda_on_host/_build_abi/libstratusgi_host.so exports the symbols
    of include/stratusgi.h.  TEST INFRASTRUCTURE: it lets the host-side state machine of csrc/sgi_abi.cu and the bodies of the GPU
    tests run where there is no GPU; nothing in the product loads it."""
    build = make_build_dir(ABI_FILES, os.path.join(HERE, "_build_abi"), whole_band_file=True)
    sanitize = "address" if sanitize is True else sanitize
    so = os.path.join(build, f"libstratusgi_host_{sanitize}.so" if sanitize else "libstratusgi_host.so")
    srcs = [os.path.join(HERE, "cuda_host_shim.h"), os.path.abspath(__file__)] + [os.path.join(CSRC, f) for f in HEADERS + ABI_FILES]
    if os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(s) for s in srcs):
        return so
    wrappers = []
    for f in ABI_FILES:
        w = os.path.join(build, "tu_" + f.replace(".cu", ".cpp"))
        text = f'// generated by host_tu.py\n#define SGI_HOST_LOCKSTEP\n#define SGI_HOST_RUNTIME\n#define SGI_HOST_TMA\n#include "cuda_host_shim.h"\n#include "{f}"\n'
        if not os.path.exists(w) or open(w).read() != text:
            open(w, "w").write(text)
        wrappers.append(w)
    # sanitize = "address": AddressSanitizer over kernels and ABI (every "device" allocation is its own malloc block, so a read or
    # write past a plane is reported); "thread": ThreadSanitizer over the lockstep runtime (two threads of a CTA touching the same
    # shared or global word without a barrier / collective in between are reported — a race check).  tools/sanitize_on_host.py;
    # the process needs LD_PRELOAD=$(gcc -print-file-name=libasan.so | libtsan.so)
    return _gxx(wrappers, so, build, (), ("-O1", "-g", "-fno-omit-frame-pointer", "-fsanitize=" + sanitize) if sanitize else ())
