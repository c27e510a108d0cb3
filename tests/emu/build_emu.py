"""Build libsph_b200_emu.so: the UNMODIFIED sources of sph_taichi_b200/csrc compiled by g++ against
tests/emu/cuda_emu.h (TEST INFRASTRUCTURE; see that header).

The only source transformation is syntactic: `kernel<<<grid, block, smem, stream>>>(args)` becomes
`emu::launch(emu::cfg(grid, block, smem, stream), kernel, args)`.

    python tests/emu/build_emu.py [--asan] [--src DIR] [--out PATH] [-D NAME=VALUE ...]
"""
import argparse
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def _match(text, i, open_ch, close_ch):
    """Index just past the bracket that closes the one at text[i]."""
    depth = 0
    for k in range(i, len(text)):
        if text[k] == open_ch:
            depth += 1
        elif text[k] == close_ch:
            depth -= 1
            if depth == 0:
                return k + 1
    raise ValueError("unbalanced " + open_ch)


def rewrite_launches(text):
    out, pos, count = [], 0, 0
    for m in re.finditer(r"<<<", text):
        if m.start() < pos:
            continue
        # kernel expression: identifier, optionally followed by <template args>, directly before <<<
        k = m.start()
        j = k
        if text[j - 1] == ">":  # template argument list
            depth, j = 0, j - 1
            while True:
                if text[j] == ">":
                    depth += 1
                elif text[j] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                j -= 1
        while j > 0 and (text[j - 1].isalnum() or text[j - 1] == "_"):
            j -= 1
        kernel = text[j:k]
        end_cfg = text.index(">>>", k)
        cfg = text[k + 3:end_cfg]
        a0 = end_cfg + 3
        while text[a0].isspace():
            a0 += 1
        assert text[a0] == "(", text[k - 40:k + 80]
        a1 = _match(text, a0, "(", ")")
        args = text[a0 + 1:a1 - 1].strip()
        out.append(text[pos:j])
        out.append(f"emu::launch(emu::cfg({cfg}), {kernel}{', ' + args if args else ''})")
        pos = a1
        count += 1
    out.append(text[pos:])
    return "".join(out), count


def build(src_dir=None, out=None, asan=False, defines=(), opt="-O1"):
    src_dir = src_dir or os.path.join(ROOT, "sph_taichi_b200", "csrc")
    build_dir = os.path.join(HERE, "_build")
    os.makedirs(build_dir, exist_ok=True)
    out = out or os.path.join(build_dir, "libsph_b200_emu_asan.so" if asan else "libsph_b200_emu.so")
    deps = [os.path.join(src_dir, f) for f in os.listdir(src_dir) if f.endswith((".cu", ".cuh"))]
    deps += [os.path.join(HERE, "cuda_emu.h"), os.path.abspath(__file__), os.path.join(ROOT, "include", "sph_b200.h")]
    if not defines and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out  # up to date (builds with extra -D always recompile)
    text = open(os.path.join(src_dir, "sph_b200.cu")).read()
    text, n = rewrite_launches(text)
    assert n > 20, f"only {n} kernel launches rewritten"
    cpp = os.path.join(build_dir, "sph_b200_emu_asan.cpp" if asan else "sph_b200_emu.cpp")
    with open(cpp, "w") as fh:
        fh.write(f'#line 1 "{os.path.join(src_dir, "sph_b200.cu")}"\n')
        fh.write(text)
    cmd = ["/usr/bin/g++", "-std=c++20", opt, "-g", "-fPIC", "-shared", "-pthread", "-ffp-contract=off",
           "-fno-strict-aliasing", "-DSPH_EMU", "-I", HERE, "-I", src_dir, "-I", os.path.join(ROOT, "include"),
           "-Wno-attributes", "-o", out, cpp]
    cmd[1:1] = [f"-D{d}" for d in defines]
    if asan:
        cmd[1:1] = ["-fsanitize=address", "-fno-omit-frame-pointer"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode:
        sys.stderr.write(res.stderr[-6000:])
        raise RuntimeError("g++ failed building the emulated library")
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--asan", action="store_true")
    ap.add_argument("--src", default=None)
    ap.add_argument("--out", default=None)
    ap.add_argument("-D", dest="defines", action="append", default=[])
    a = ap.parse_args()
    print(build(a.src, a.out, a.asan, a.defines))
