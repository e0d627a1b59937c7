#!/usr/bin/env python
"""xmake -- the repo's own build tool, same target files as the reference's (`xmake.yml` next to the
sources: target -> {main, srcs, hdrs, deps, rule, lopts, gopts}; reference build/xmake.cc:92-103,
280-312), with the rule the reference left unused filled in:

    rule: c++    g++ -O3 -std=c++20 (reference flags minus -mavx -ffast-math: the host only
                 orchestrates, the arithmetic is on the device)
    rule: cuda   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo, objects linked into one
                 shared library (lopts: ["-shared"]) that C++ targets depend on through its
                 extern "C" header only

Usage (from anywhere inside the repo):  python tools/xmake.py //dependence_free_rl_b200/host/apps/bin_packing/ppo_training
Targets are `//path/from/repo/root/name`; `deps` use the same form. Rebuilds are mtime based.
Outputs go to `<package dir>/.out/`.
"""
import os
import subprocess
import sys

import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CXX = os.environ.get("CXX", "g++")
NVCC = os.environ.get("NVCC", "nvcc")
CXXFLAGS = ["-O3", "-std=c++20", "-Wall", "-I" + os.path.join(ROOT, "dependence_free_rl_b200", "host"),
            "-I" + os.path.join(ROOT, "include")]
NVFLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def load(pkg):
    path = os.path.join(ROOT, pkg, "xmake.yml")
    with open(path) as f:
        return yaml.safe_load(f) or {}


def split(label):
    assert label.startswith("//"), f"target labels look like //pkg/name, got {label}"
    pkg, name = label[2:].rsplit("/", 1)
    return pkg, name


def mtime(p):
    return os.path.getmtime(p) if os.path.exists(p) else 0.0


def run(cmd):
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build(label, done):
    """Returns (objects, libraries, newest input mtime) contributed by `label`."""
    if label in done:
        return done[label]
    pkg, name = split(label)
    t = load(pkg).get(name)
    if t is None:
        raise SystemExit(f"xmake: no target {name} in {pkg}/xmake.yml")
    pdir = os.path.join(ROOT, pkg)
    out = os.path.join(pdir, ".out")
    os.makedirs(out, exist_ok=True)
    rule = t.get("rule", "c++")
    objs, libs, newest = [], [], 0.0
    for d in t.get("deps", []) or []:
        o, l, m = build(d, done)
        objs += o
        libs += l
        newest = max(newest, m)
    hdr_m = max([mtime(os.path.join(pdir, h)) for h in t.get("hdrs", []) or []] + [newest])
    my_objs = []
    for s in t.get("srcs", []) or []:
        src = os.path.join(pdir, s)
        obj = os.path.join(out, os.path.splitext(s)[0] + ".o")
        if mtime(obj) < max(mtime(src), hdr_m):
            if rule == "cuda":
                run([NVCC] + NVFLAGS + (t.get("gopts") or []) + ["-c", src, "-o", obj])
            else:
                run([CXX] + CXXFLAGS + (t.get("gopts") or []) + ["-c", src, "-o", obj])
        my_objs.append(obj)
        hdr_m = max(hdr_m, mtime(obj))
    lopts = t.get("lopts") or []
    if rule == "cuda" and "-shared" in lopts:
        lib = os.path.join(out, "lib" + name + ".so")
        if mtime(lib) < hdr_m:
            run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib] + my_objs + ["-ldl"])
        res = ([], libs + [lib], max(hdr_m, mtime(lib)))
    elif t.get("main"):
        exe = os.path.join(out, name)
        if mtime(exe) < hdr_m:
            link = []
            for l in dict.fromkeys(libs):
                link += ["-L" + os.path.dirname(l), "-l" + os.path.basename(l)[3:-3], "-Wl,-rpath," + os.path.dirname(l)]
            run([CXX] + my_objs + objs + link + lopts + ["-pthread", "-o", exe])
        res = ([], [], mtime(exe))
    else:
        res = (objs + my_objs, libs, hdr_m)
    done[label] = res
    return res


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    for label in sys.argv[1:]:
        if not label.startswith("//"):  # bare name: a target of the package in the current directory
            label = "//" + os.path.relpath(os.getcwd(), ROOT) + "/" + label
        build(label, {})


if __name__ == "__main__":
    main()
