"""TEST INFRASTRUCTURE: spot-check of library combinations that are not in the test suite (csrc/Makefile `custom`).

For each combination of the reference's compile-time switches below, the kernel sources are built for the SIMT
interpreter with the product's defines, the REFERENCE is built the same way from a scratch copy (the switches are
unguarded #defines there), and both are driven through the drop-in interface: every state array must be identical.
Needs /root/reference.      python tests/simt/check_combinations.py
"""
import sys, os, subprocess, tempfile, shutil, dataclasses
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.chdir(ROOT)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "simt"))
import support as S, build as B
from ntsc_crt_b200 import layout
import numpy as np
REF = "/root/reference"
combos = {  # name: (base variant, product defs, reference sed edits {file: [(old, new)]}, reference sources, CRT_SYSTEM)
  "template_conv": ("template", ["-DCRT_SYSTEM=4", "-DCRTX_CONV=1"], {"crt_core.c": [("#define USE_CONVOLUTION 0", "#define USE_CONVOLUTION 1")]}, ["crt_core.c", "crt_template.c"], 4),
  "snes_conv": ("snes", ["-DCRT_SYSTEM=3", "-DCRTX_CONV=1"], {"crt_core.c": [("#define USE_CONVOLUTION 0", "#define USE_CONVOLUTION 1")]}, ["crt_core.c", "crt_snes.c"], 3),
  "pv1k_bloom": ("pv1k", ["-DCRT_SYSTEM=2", "-DCRT_DO_BLOOM=1"], {"crt_core.h": [("#define CRT_DO_BLOOM    0", "#define CRT_DO_BLOOM    1")]}, ["crt_core.c", "crt_pv1k.c"], 2),
  "snes_bloom": ("snes", ["-DCRT_SYSTEM=3", "-DCRT_DO_BLOOM=1"], {"crt_core.h": [("#define CRT_DO_BLOOM    0", "#define CRT_DO_BLOOM    1")]}, ["crt_core.c", "crt_snes.c"], 3),
  "template_bloom": ("template", ["-DCRT_SYSTEM=4", "-DCRT_DO_BLOOM=1"], {"crt_core.h": [("#define CRT_DO_BLOOM    0", "#define CRT_DO_BLOOM    1")]}, ["crt_core.c", "crt_template.c"], 4),
}
srcdir = os.path.join(B.OUT, "src"); B.prepare(srcdir)
for name, (base, defs, edits, srcs, sysn) in combos.items():
    lib = os.path.join(B.OUT, "combo_libcrt_simt_%s.so" % name)
    cmd = ["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-fwrapv", "-fno-strict-aliasing", "-Wno-unknown-pragmas", "-Wno-attributes", "-I" + srcdir, "-I" + os.path.join(B.ROOT, "include")] + defs + ["-o", lib] + [os.path.join(srcdir, f) for f in ("crtx.cpp", "crt_dropin.cpp", "simt_runtime.cpp")]
    subprocess.run(cmd, check=True)
    T = tempfile.mkdtemp()
    for f in os.listdir(REF):
        if f.endswith((".c", ".h")): shutil.copy(os.path.join(REF, f), T)
    for f, reps in edits.items():
        txt = open(os.path.join(T, f)).read()
        for a, b in reps:
            assert a in txt, (f, a); txt = txt.replace(a, b, 1)  # the switch itself, not any guard further down
        open(os.path.join(T, f), "w").write(txt)
    reflib = os.path.join(B.OUT, "combo_libref_%s.so" % name)
    subprocess.run(["gcc", "-O3", "-fPIC", "-w", "-shared", "-DCRT_SYSTEM=%d" % sysn, "-I" + T, "-o", reflib] + [os.path.join(T, f) for f in srcs] + ["oracle/ref_shim.c"], check=True)
    shutil.rmtree(T)
    layout.SPECS[name] = dataclasses.replace(layout.SPECS[base], name=name)
    ok = True
    for (outw, outh, fmt) in [(640, 480, 5), (333, 250, 0)]:
        img = S.rand_image(300, 260, seed=3)
        gpu = S.CEngine(lib, name, outw, outh, fmt); ref = S.CEngine(reflib, name, outw, outh, fmt)
        for e in (gpu, ref):
            e.set(blend=1, scanlines=1)
            for it in range(4):
                e.modulate(img, format=5, as_color=1, field=it & 1, frame=0, dot_crawl_offset=it, hue=20 * it)
                e.demodulate(5 * it)
        try:
            S.assert_same_state(gpu.state(), ref.state(), name)
        except AssertionError as ex:
            ok = False; print(name, "MISMATCH", str(ex)[:200])
    print(name, "identical to the reference built the same way" if ok else "FAILED")
