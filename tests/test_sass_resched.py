"""CPU-side checks of the experiment tool tools/sass_resched.py (post-pass SASS scheduler; profiles/r02_experiments.md):
on a small butterfly kernel compiled here for sm_100a the re-ordered cubin must hold exactly the same instructions (a
permutation inside the kernel, control fields aside), the tool's own symbolic equivalence check must pass, every
fixed-latency dependence must keep ptxas' minimum distance, and the control-field-only modes must not move anything."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "tools", "sass_resched.py")

SRC = r"""
#include <cstdint>
__device__ __forceinline__ uint32_t csub(uint32_t x, uint32_t negm) { return __viaddmin_u32(x, negm, x); }
extern "C" __global__ void __launch_bounds__(64) bfly(uint32_t *d, const uint2 *tw, uint32_t negq, uint32_t neg2q,
                                                       uint32_t twoq, uint32_t zero) {
    uint32_t x[32];
#pragma unroll
    for (int k = 0; k < 32; k++) x[k] = d[threadIdx.x + 64 * k];
#pragma unroll
    for (int j = 0; j < 5; j++) {
        const int half = 16 >> j;
#pragma unroll
        for (int g = 0; g < (1 << j); g++) {
            const uint2 w = tw[(1 << j) + g];
#pragma unroll
            for (int i = 0; i < half; i++) {
                uint32_t &a = x[g * 2 * half + i], &b = x[g * 2 * half + i + half];
                const uint32_t t = csub(a, neg2q);
                const uint32_t Q = b * w.x + __umulhi(b, w.y) * negq;
                a = t + Q + zero;
                b = t + twoq - Q;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 32; k++) d[threadIdx.x + 64 * k] = x[k];
}
"""

pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None or shutil.which("cuobjdump") is None,
                                reason="needs nvcc and cuobjdump")


@pytest.fixture(scope="module")
def cubin(tmp_path_factory):
    d = tmp_path_factory.mktemp("resched")
    src = d / "bfly.cu"
    src.write_text(SRC)
    out = d / "bfly.cubin"
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-cubin", "-o", str(out), str(src)],
                   check=True)
    return out


def _words(path):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_resched as S
    sass = subprocess.run(["cuobjdump", "-sass", str(path)], capture_output=True, text=True, check=True).stdout
    part = re.split(r"\n\s*Function : ", sass)[1]
    return S, S.parse_function(part)


def _run(cubin, out, *args):
    r = subprocess.run([sys.executable, TOOL, str(cubin), str(out), "--kernel", "bfly", "--report", "--min-movable", "32",
                        *args], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


CTRL = (0xF << 41) | (1 << 45) | (0xF << 58)          # stall, yield, reuse: what the tool may rewrite


def test_reorder_is_a_permutation_with_legal_distances(cubin, tmp_path):
    out = tmp_path / "rs.cubin"
    log = _run(cubin, out, "--reuse", "clear")
    assert "instructions rewritten" in log
    S, a = _words(cubin)
    _, b = _words(out)
    assert len(a) == len(b)
    key = lambda x: (x.lo, x.hi & ~CTRL)
    assert sorted(map(key, a)) == sorted(map(key, b)), "not a permutation of the original instructions"
    moved = sum(1 for x, y in zip(a, b) if key(x) != key(y))
    assert moved > 50, "nothing was re-ordered"
    # scoreboard fields travel with their instruction; fences stay where they were
    for x, y in zip(a, b):
        if x.kind == "fence":
            assert key(x) == key(y), "a fence moved: %s" % x.text
    # fixed-latency RAW distances in the new order, from the stall counts the tool wrote
    t, issue = 0, {}
    last_writer = {}
    for y in b:
        for r in y.src:
            if r in last_writer:
                p, tp = last_writer[r]
                if p.kind == "alu":
                    need = S.LAT_SAME if (y.kind == "alu" and y.pipe == p.pipe) else S.LAT_CROSS
                    assert t - tp >= need, "%s issued %d cycles after %s" % (y.text, t - tp, p.text)
        for r in y.dst:
            last_writer[r] = (y, t)
        t += max(y.stall, 1)
    # and the file differs only inside the kernel's .text section
    A, B = open(cubin, "rb").read(), open(out, "rb").read()
    assert len(A) == len(B)
    secs = S.elf_sections(bytearray(A))
    off, size = secs[".text.bfly"]
    assert A[:off] == B[:off] and A[off + size:] == B[off + size:]


def test_control_only_modes_move_nothing(cubin, tmp_path):
    for args in (("--order", "keep", "--reuse", "clear"), ("--order", "keep", "--reuse", "keep", "--yield-policy", "all1")):
        out = tmp_path / "k.cubin"
        _run(cubin, out, *args)
        _, a = _words(cubin)
        _, b = _words(out)
        for x, y in zip(a, b):
            assert (x.lo, x.hi & ~CTRL) == (y.lo, y.hi & ~CTRL)
            assert x.stall == y.stall
        if "clear" in args:
            assert all(((y.hi >> 58) & 0xF) == 0 for y in b if y.kind == "alu")


def test_model_reproduces_ptxas_stall_total(cubin, tmp_path):
    """ptxas' own order re-timed by the tool's latency model must not need more cycles than ptxas gave it"""
    log = _run(cubin, tmp_path / "p.cubin", "--order", "keep", "--restall", "--reuse", "keep")
    m = re.search(r"single-warp issue cycles (\d+) -> (\d+)", log)
    assert m, log
    before, after = int(m.group(1)), int(m.group(2))
    assert after <= before and after >= 0.9 * before, (before, after)
