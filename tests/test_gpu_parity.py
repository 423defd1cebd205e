"""Parity of the CUDA path (through the C ABI) against the reference's golden vectors and the CPU
oracle, on a real B200.  Same cases as tests/test_emu_kernels.py, device = cuda:0."""
import numpy as np
import pytest
import torch

import kernel_cases as kc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def real_library():
    from fumi_b200 import _lib
    from fumi_b200 import build
    L = _lib.lib()
    assert L._name == build.LIB, "GPU tests must run on the CUDA build of libfumi_b200.so"
    assert L.fumi_device_sm_count() > 0
    yield


def test_dense():
    kc.dense_case(DEV)


def test_gemm_tcgen05_3xtf32():
    kc.gemm_tc_case(DEV)


def test_gemm_tcgen05_3xf16():
    kc.gemm_f16_case(DEV)


@pytest.mark.parametrize("precision", [1, 2])
@pytest.mark.parametrize("via", ["dict", "bank"])
def test_fumi_train_tensor_core_dense(via, precision):
    kc.fumi_train_case(DEV, "fumi_train_n5k5_d512", via=via, precision=precision)


def test_warp_gemm_f16_planes():
    kc.warp_gemm_f16_case(DEV)


def test_maml_test_then_train_keeps_flat_gradients():
    kc.maml_test_then_train_case(DEV)


def test_gram():
    kc.gram_case(DEV, big=True)


def test_adam():
    kc.adam_case(DEV)


@pytest.mark.parametrize("via", ["dict", "bank"])
def test_fumi_train_n5k5(via):
    kc.fumi_train_case(DEV, "fumi_train_n5k5_d512", via=via)


@pytest.mark.parametrize("precision", [0, 2])
def test_fumi_train_tanh(precision):
    kc.fumi_train_case(DEV, "fumi_train_n5k5_d512_tanh", precision=precision)


@pytest.mark.parametrize("precision", [0, 2])
def test_fumi_train_n20k5_multitile(precision):
    kc.fumi_train_case(DEV, "fumi_train_n20k5_d512", precision=precision)


def test_benched_config_gradients_vs_fp64_oracle():
    """BASELINE configs[1] as benched: D = 2048, T = 768, NQ = 160, precision 2, dropout 0.25 -- all 8 meta-gradients
    (dW0 included) against the fp64 oracle."""
    par = kc.oracle_train_case(DEV, N=5, K=5, Q=32, steps=5, D=2048, T=768, B=8, dropout=0.25, precision=2)
    print("benched-config parity:", {k: v for k, v in par.items() if k != "grad_relerr"})


def test_config5_shape_gradients_vs_fp64_oracle():
    """BASELINE configs[4] shape: 20-way 5-shot, NK = 100, NQ = 640, 10 inner steps, precision 2."""
    par = kc.oracle_train_case(DEV, N=20, K=5, Q=32, steps=10, D=512, T=64, B=2, dropout=0.25, precision=2)
    print("config-5 parity:", {k: v for k, v in par.items() if k != "grad_relerr"})


def test_one_shot_and_ten_way_gradients_vs_fp64_oracle():
    """The other class-count buckets / tile shapes of the tensor-core kernels: 5-way 1-shot (NK = 5, one 16-row tile),
    10-way 3-shot (NK = 30, N = 10)."""
    kc.oracle_train_case(DEV, N=5, K=1, Q=8, steps=5, D=512, T=64, B=3, dropout=0.25, precision=2)
    kc.oracle_train_case(DEV, N=10, K=3, Q=6, steps=3, D=512, T=64, B=3, dropout=0.0, precision=2, tanh=True)
    kc.oracle_train_case(DEV, N=7, K=4, Q=5, steps=2, D=512, T=64, B=2, dropout=0.1, precision=2)


def test_fumi_train_many_classes_few_rows():
    """NK <= 32 with more than 11 classes (16-way 2-shot: NK = 32): outside the fp16-plane kernels' class buckets, this
    lands on round 1's tensor-core kernels in normal operation (episode.cu, path 1) -- same oracle, same bar."""
    kc.oracle_train_case(DEV, N=16, K=2, Q=4, steps=3, D=512, T=64, B=3, dropout=0.25, precision=2)
    kc.oracle_train_case(DEV, N=12, K=1, Q=3, steps=2, D=512, T=64, B=2, dropout=0.0, precision=2)


def test_fumi_evaluate_api():
    kc.fumi_evaluate_api_case(DEV, "fumi_train_n5k5_d512")


@pytest.mark.parametrize("name", ["fumi_test_n5k1_full", "fumi_test_n5k5_full"])
@pytest.mark.parametrize("via", ["dict", "bank"])
def test_fumi_meta_test_100_steps(name, via):
    n_ties = kc.fumi_test_case(DEV, name, via=via)
    print(f"{name}/{via}: {n_ties} ReLU gate ties (|z| < 1e-5) over the 100-step unroll")
    assert n_ties <= 4, "more ReLU-gate ties than rounding noise explains"


@pytest.mark.parametrize("name", ["maml_train_n5k5_d512", "maml_train_n5k5_d512_fo", "maml_test_n5k5_d512"])
def test_maml(name):
    kc.maml_case(DEV, name)


def test_am3():
    kc.am3_case(DEV)


def test_am3_train_step():
    kc.am3_train_case(DEV)


def test_dropout_masks():
    kc.dropout_case(DEV)


@pytest.mark.parametrize("name,N,K,Qtrain", [("sampler_n5k5b4", 5, 5, 32), ("sampler_n5k1b3", 5, 1, 32),
                                              ("sampler_n10k5b2", 10, 5, 20), ("sampler_n20k5b2", 20, 5, 16)])
def test_device_sampler_golden(name, N, K, Qtrain):
    kc.device_sampler_golden_case("cuda", name, N, K, Qtrain)


def test_device_sampler_equals_host_sampler():
    kc.device_sampler_vs_host_case("cuda", B=64)


def test_device_loader_prefetch():
    kc.device_loader_case("cuda")


def test_no_cpu_path():
    """The product refuses non-CUDA devices instead of falling back."""
    from fumi_b200 import _lib
    from fumi_b200.engine import EpisodeEngine
    with pytest.raises(_lib.FumiError):
        EpisodeEngine("cpu")
