"""TEST INFRASTRUCTURE: build tests/simt/_build/libcrt_simt_<variant>.so -- the product's CUDA sources compiled
by g++ against the SIMT interpreter in this directory (cuda_runtime.h, crt_ptx.cuh, simt_runtime.cpp), so kernel
logic can be executed and checked against the oracle where there is no GPU.

The sources are used as they are except for two mechanical rewrites that plain C++ needs:
  * `kernel<<<grid, block, smem, stream>>>(args)`  ->  `::simt::launch(grid, block, smem, [&]{ kernel(args); })`
  * `extern __shared__ T name[]`                    ->  `extern T name[]` (defined in simt_runtime.cpp)
and crt_ptx.cuh (the inline PTX) is replaced by its stand-in here.  Nothing in the product uses these libraries.
"""
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "ntsc-crt_b200", "csrc")
OUT = os.path.join(HERE, "_build")


def variant_defines():
    """DEF_<variant> lines of the product Makefile -- one source of truth for what a variant is."""
    defs = {}
    for line in open(os.path.join(CSRC, "Makefile")):
        m = re.match(r"DEF_(\w+)\s*:=\s*(.*)", line)
        if m:
            defs[m.group(1)] = m.group(2).split()
    return defs


def _balanced_back(text, pos):
    """text[pos] == '>' closing a template argument list: index of the matching '<'."""
    depth = 0
    i = pos
    while i >= 0:
        c = text[i]
        if c == ">":
            depth += 1
        elif c == "<":
            depth -= 1
            if depth == 0:
                return i
        i -= 1
    raise ValueError("unbalanced template arguments")


def _split_top(s):
    parts, depth, cur = [], 0, ""
    for c in s:
        if c in "(<[{":
            depth += 1
        elif c in ")>]}":
            depth -= 1
        if c == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += c
    parts.append(cur.strip())
    return parts


def rewrite_launches(text):
    out = ""
    while True:
        k = text.find("<<<")
        if k < 0:
            return out + text
        # kernel name (with optional template arguments) before <<<
        i = k - 1
        while text[i].isspace():
            i -= 1
        if text[i] == ">":
            i = _balanced_back(text, i) - 1
        while text[i].isalnum() or text[i] in "_:":
            i -= 1
        name_start = i + 1
        name = text[name_start:k].strip()
        e = text.index(">>>", k)
        cfg = _split_top(text[k + 3:e])
        while len(cfg) < 3:
            cfg.append("0")
        # argument list
        a0 = text.index("(", e)
        depth, j = 0, a0
        while True:
            if text[j] == "(":
                depth += 1
            elif text[j] == ")":
                depth -= 1
                if depth == 0:
                    break
            j += 1
        args = text[a0 + 1:j]
        out += text[:name_start]
        out += "::simt::launch(dim3(%s), dim3(%s), (size_t) (%s), [&]() { %s(%s); })" % (cfg[0], cfg[1], cfg[2], name, args)
        text = text[j + 1:]


def prepare(dst):
    os.makedirs(dst, exist_ok=True)
    for f in sorted(os.listdir(CSRC)):
        if not f.endswith((".cu", ".cuh", ".h")) or f == "crt_ptx.cuh":
            continue
        src = open(os.path.join(CSRC, f)).read()
        src = rewrite_launches(src)
        src = re.sub(r"extern\s+__shared__", "extern", src)
        name = f[:-3] + ".cpp" if f.endswith(".cu") else f
        with open(os.path.join(dst, name), "w") as o:
            o.write(src)
    for f in ("cuda_runtime.h", "crt_ptx.cuh", "simt_runtime.cpp"):
        shutil.copy(os.path.join(HERE, f), os.path.join(dst, f))


def lib_path(variant):
    return os.path.join(OUT, "libcrt_simt_%s.so" % variant)


def build(variants=None, force=False, asan=False, align=True):
    """asan=True: AddressSanitizer build into _build/asan/ (run python with LD_PRELOAD=$(gcc -print-file-name=libasan.so),
    ASAN_OPTIONS=detect_leaks=0 and SIMT_TIGHT_ALLOC=1): out-of-bounds accesses of "device" and "shared" memory."""
    defs = variant_defines()
    variants = list(variants) if variants else sorted(defs)
    srcdir = os.path.join(OUT, "src")
    prepare(srcdir)
    newest = max(os.path.getmtime(os.path.join(d, f)) for d in (CSRC, HERE, os.path.join(ROOT, "include"))
                 for f in os.listdir(d) if os.path.isfile(os.path.join(d, f)))
    procs = []
    if asan:
        os.makedirs(os.path.join(OUT, "asan"), exist_ok=True)
    for v in variants:
        lib = lib_path(v) if not asan else os.path.join(OUT, "asan", os.path.basename(lib_path(v)))
        if not force and os.path.exists(lib) and os.path.getmtime(lib) >= newest:
            continue
        cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-fwrapv", "-fno-strict-aliasing",
               "-Wno-unknown-pragmas", "-Wno-attributes", "-I" + srcdir, "-I" + os.path.join(ROOT, "include")]
        if asan:
            cmd += ["-fsanitize=address", "-fno-omit-frame-pointer"]
        if align:  # every dereference is checked against its type's alignment (short 2, uint2 8, uint4 16, as on the GPU,
            # where a misaligned access is a fatal "misaligned address"); a violation traps (SIGILL).  To see where:
            # build with -fsanitize=alignment alone and LD_PRELOAD=$(gcc -print-file-name=libubsan.so)
            cmd += ["-fsanitize=alignment", "-fsanitize-undefined-trap-on-error"]
        cmd += defs[v] + ["-o", lib] + [os.path.join(srcdir, f) for f in ("crtx.cpp", "crt_dropin.cpp", "simt_runtime.cpp")]
        procs.append((v, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for v, p in procs:
        log = p.communicate()[0]
        if p.returncode:
            sys.stderr.write(log)
            raise RuntimeError("simt build of %s failed" % v)
    return [lib_path(v) if not asan else os.path.join(OUT, "asan", os.path.basename(lib_path(v))) for v in variants]


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if a not in ("--asan", "--no-align")]
    print("\n".join(build(args or None, force=True, asan="--asan" in sys.argv, align="--no-align" not in sys.argv)))
