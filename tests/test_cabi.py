"""CPU-side checks of the C ABI: the library loads, exports every symbol include/usac_gpu.h declares, and refuses to
work (loudly, no CPU fallback) when there is no CUDA device. No compute calls here."""
import ctypes as C
import os
import subprocess

import pytest

from ransac_b200 import capi


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(capi.LIB_PATH):
        from ransac_b200 import build
        build.build()
    return capi.load()


def test_exports_every_declared_symbol(lib):
    names = capi.declared_symbols()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), n
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert set(names) <= exported


def test_struct_layouts_match_header():
    # usac_sampler_cfg / usac_fit_cfg / usac_fit_result as laid out by a C compiler for the same header
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "usac_gpu.h"
    int main(void) {
        printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(usac_sampler_cfg), sizeof(usac_fit_cfg), sizeof(usac_fit_result),
               offsetof(usac_fit_cfg, threshold), offsetof(usac_fit_cfg, sample_table), offsetof(usac_fit_result, best_hyp),
               offsetof(usac_fit_result, evals));
        return 0;
    }'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.dirname(capi.HEADER_PATH), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        vals = list(map(int, subprocess.check_output([os.path.join(d, "t")]).split()))
    assert vals == [C.sizeof(capi.SamplerCfg), C.sizeof(capi.FitCfg), C.sizeof(capi.FitResult), capi.FitCfg.threshold.offset,
                    capi.FitCfg.sample_table.offset, capi.FitResult.best_hyp.offset, capi.FitResult.evals.offset]


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = lib.usac_gpu_create(C.byref(h), 0)
    assert rc == capi.ERR_CUDA and not h.value
    assert b"no CPU fallback" in lib.usac_gpu_last_error(None)


def test_product_never_imports_oracle():
    root = os.path.join(os.path.dirname(capi.__file__))
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower().replace("no cpu fallback", ""), os.path.join(dirpath, f)
