"""Kernel bodies on the host emulation (tests/emu): the same parity cases the GPU tests run.

TEST INFRASTRUCTURE: the emulation library is libfumi_b200's own kernel source compiled for the CPU
with a CUDA shim; it checks indexing / synchronisation / math where no GPU exists.  The GPU
results are what counts (tests/test_gpu_parity.py, -m gpu)."""
import os
import sys

import pytest

from fumi_b200 import _lib

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import kernel_cases as kc  # noqa: E402


# The full emulated set takes ~6 min on 8 cores; the default CPU run keeps one case per kernel path and the
# rest (duplicates of what the GPU suite runs natively) is enabled with FUMI_EMU_FULL=1.
full = pytest.mark.skipif(os.environ.get("FUMI_EMU_FULL") != "1", reason="set FUMI_EMU_FULL=1 for the full emulated set")


@pytest.fixture(scope="module", autouse=True)
def emu_lib():
    import inject
    with inject.emulation():
        yield


def test_dense():
    kc.dense_case("cpu")


def test_warp_gemm_f16_planes():
    kc.warp_gemm_f16_case("cpu")


def test_gram():
    kc.gram_case("cpu")


def test_adam():
    kc.adam_case("cpu")


def test_maml_test_then_train_keeps_flat_gradients():
    kc.maml_test_then_train_case("cpu")


@pytest.mark.parametrize("via", ["dict", pytest.param("bank", marks=full)])
def test_fumi_train_n5k5(via):
    kc.fumi_train_case("cpu", "fumi_train_n5k5_d512", via=via)


@full
def test_fumi_train_tanh():
    kc.fumi_train_case("cpu", "fumi_train_n5k5_d512_tanh")


def test_fumi_train_n20k5_multitile():
    kc.fumi_train_case("cpu", "fumi_train_n20k5_d512")


@full
def test_fumi_evaluate_api():
    kc.fumi_evaluate_api_case("cpu", "fumi_train_n5k5_d512")


@full
def test_fumi_test_n5k1_100_steps():
    kc.fumi_test_case("cpu", "fumi_test_n5k1_full")


@pytest.mark.parametrize("name", [pytest.param("maml_train_n5k5_d512", marks=full), "maml_train_n5k5_d512_fo",
                                  "maml_test_n5k5_d512"])
def test_maml(name):
    kc.maml_case("cpu", name)


def test_am3():
    kc.am3_case("cpu")


def test_dropout_masks():
    kc.dropout_case("cpu")


@pytest.mark.parametrize("name,N,K,Qtrain", [("sampler_n5k5b4", 5, 5, 32), pytest.param("sampler_n20k5b2", 20, 5, 16, marks=full)])
def test_device_sampler_golden(name, N, K, Qtrain):
    kc.device_sampler_golden_case("cpu", name, N, K, Qtrain)


def test_device_sampler_equals_host_sampler():
    kc.device_sampler_vs_host_case("cpu", B=6)


def test_device_loader_prefetch():
    kc.device_loader_case("cpu")


def test_am3_train_step():
    kc.am3_train_case("cpu")

