"""Every env-switched kernel variant of the library, proven on the GPU box: the switches are read once per process
(function-local statics in gemm.cu / gemv.cu / attention*.cu), so each variant runs the relevant parity tests of
tests/test_gpu_parity.py in its OWN subprocess with the switch set.  All subprocesses are started together (they are
small 2-layer-geometry tests) and each parametrised test below asserts on one of them.
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

GEMM_K = "gemm_tcgen05 or gemm_linearity or forward_tiny or prefix_reuse"
ATTN_K = "omni_attention or forward_tiny or forward_masks or t2i_512 or prefix_reuse"
DECODE_K = "skinny or mmu_generate_batched or step_and_decode or megakernel"
VARIANTS = {
    "attention_tcgen05": ({"SHOWO_ATTN_TC": "1"}, ATTN_K),
    "attention_mma_sync": ({"SHOWO_ATTN_TC": "0"}, ATTN_K),
    "gemm_streamk": ({"SHOWO_GEMM_STREAMK": "1"}, GEMM_K + " or streamk"),
    "gemm_order_m_first": ({"SHOWO_GEMM_ORDER": "0"}, GEMM_K + " or token_major"),
    "gemm_l2_evict_last_hint": ({"SHOWO_GEMM_HINT": "1"}, GEMM_K + " or token_major"),
    "gemm_one_cta": ({"SHOWO_GEMM_CG": "1"}, GEMM_K),
    "gemm_cluster_multicast": ({"SHOWO_GEMM_CG": "1", "SHOWO_GEMM_CL": "2"}, GEMM_K),
    "gemm_bk32": ({"SHOWO_GEMM_CG": "1", "SHOWO_GEMM_BK": "32"}, GEMM_K),
    "gemm_pair_bk64": ({"SHOWO_GEMM_BK": "64"}, GEMM_K),
    "skinny_register_prefetch": ({"SHOWO_SKINNY": "1"}, DECODE_K),
    "decode_attention_per_thread": ({"SHOWO_DECODE_ATTN": "1"}, DECODE_K),
    "decode_ln_fused": ({"SHOWO_DECODE_LN_FUSED": "1"}, DECODE_K),
    "decode_l2_prefetch": ({"SHOWO_L2_PREFETCH": "1"}, DECODE_K),
    "wgrad_transposed_copies": ({"SHOWO_WGRAD_MN": "0"}, None),
    "ln_fold_off": ({"SHOWO_LN_FOLD": "0"}, DECODE_K),
    "ln_fold_all_paths": ({"SHOWO_LN_FOLD": "2"}, GEMM_K + " or forward_masks or t2i or " + DECODE_K),
    "no_pdl": ({"SHOWO_PDL": "0"}, "forward_tiny or mmu_generate_batched or teacher_forced"),
}


@pytest.fixture(scope="module")
def runs():
    procs = {}
    for name, (env, k) in VARIANTS.items():
        e = dict(os.environ)
        e.update(env)
        # k = None: the variant concerns the training step -> tests/test_gpu_train.py as a whole
        cmd = [sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py" if k else "test_gpu_train.py"), "-q", "-x",
               "-m", "gpu", "-p", "no:cacheprovider"] + (["-k", k] if k else [])
        procs[name] = subprocess.Popen(cmd, cwd=ROOT, env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    out = {}
    for name, p in procs.items():
        try:
            txt, _ = p.communicate(timeout=900)
        except subprocess.TimeoutExpired:
            p.kill()
            txt, _ = p.communicate()
            txt += "\n[timeout]"
        out[name] = (p.returncode, txt)
    return out


@pytest.mark.parametrize("name", list(VARIANTS))
def test_variant(runs, name):
    rc, txt = runs[name]
    tail = "\n".join(txt.strip().splitlines()[-15:])
    print(f"{name} {VARIANTS[name][0]}: {txt.strip().splitlines()[-1] if txt.strip() else ''}")
    assert rc == 0, f"variant {name} {VARIANTS[name][0]} failed:\n{tail}"
    assert " passed" in txt and "no tests ran" not in txt, tail
