"""Golden fixtures written by the upstream reference itself (oracle/make_golden.py) vs the oracle, the emulated
kernels (CPU) and the CUDA path (-m gpu)."""
import pytest

import common
import golden_util as gu
from emu_harness import emulated
from mdn_sfm_b200 import synthetic


@pytest.mark.parametrize("mode", list(gu.NETINIT_MODES))
def test_oracle_reproduces_netinit_golden(mode):
    photo, ssim_on, dmin = gu.NETINIT_MODES[mode]
    (B, H, W), batch = gu.load_netinit_batch()
    opt = synthetic.default_opt(B, H, W, disable_min=dmin)
    got = common.oracle_run(opt, batch, mode, photo, ssim_on, "cpu", pose_grad=True, rule_ties=False)
    gu.check_against_golden(gu.load_outputs("netinit_" + mode), got, photo, 1e-6, 1e-5)


@pytest.mark.parametrize("mode", list(gu.NETINIT_MODES))
def test_emulated_kernels_reproduce_netinit_golden(mode):
    photo, ssim_on, dmin = gu.NETINIT_MODES[mode]
    (B, H, W), batch = gu.load_netinit_batch()
    opt = synthetic.default_opt(B, H, W, disable_min=dmin)
    with emulated():
        got = common.product_run(opt, batch, mode, photo, ssim_on, "cpu", pose_grad=True)
        gu.check_against_golden(gu.load_outputs("netinit_" + mode), got, photo, common.FWD_TOL, common.GRAD_TOL)


def test_oracle_reproduces_one_stress_golden():
    # the other seven stress fixtures are covered on the GPU; one here keeps the CPU suite short
    photo, ssim_on, dmin = gu.STRESS_MODES["T"]
    z = gu.load_outputs("stress_T_seed42")
    opt, batch = common.make(4, 192, 640, seed=int(z["seed"]), flow_std=float(z["flow_std"]))
    got = common.oracle_run(opt, batch, "T", photo, ssim_on, "cpu", pose_grad=True)
    gu.check_against_golden(z, got, photo, 1e-6, 1e-5, full=False)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", list(gu.NETINIT_MODES))
def test_cuda_reproduces_netinit_golden(mode):
    photo, ssim_on, dmin = gu.NETINIT_MODES[mode]
    (B, H, W), batch = gu.load_netinit_batch()
    opt = synthetic.default_opt(B, H, W, disable_min=dmin)
    z = gu.load_outputs("netinit_" + mode)
    got = common.product_run(opt, batch, mode, photo, ssim_on, "cuda", pose_grad=True, arith="cpu")  # fixtures come from the CPU run of the reference
    # This fixture has a near-degenerate pose (random-init PoseNet: |t| ~ 1e-4), so p2^T F p1 cancels almost completely
    # and the reference's OWN arithmetic, run eagerly on this GPU, differs from its CPU run (the fixture) by ~1-2e-5 in the
    # epipolar maps (cuBLAS vs MKL rounding of F; profiles/ has the measured table).  The maps are therefore held to
    # max(1e-5, 1.5 x that measured reference-vs-reference floor); scalars and gradients keep the plain tolerances.
    floor = gu.map_noise_floor(z, common.oracle_run(opt, batch, mode, photo, ssim_on, "cuda", rule_ties=False))
    gu.check_against_golden(z, got, photo, common.FWD_TOL, common.GRAD_TOL, map_tol=max(common.FWD_TOL, 1.5 * floor))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", list(gu.STRESS_MODES))
@pytest.mark.parametrize("seed", [42, 43])
def test_cuda_reproduces_stress_golden(mode, seed):
    photo, ssim_on, dmin = gu.STRESS_MODES[mode]
    z = gu.load_outputs("stress_%s_seed%d" % (mode, seed))
    opt, batch = common.make(4, 192, 640, seed=int(z["seed"]), flow_std=float(z["flow_std"]))
    got = common.product_run(opt, batch, mode, photo, ssim_on, "cuda", pose_grad=True, arith="cpu")  # fixtures come from the CPU run of the reference
    gu.check_against_golden(z, got, photo, common.FWD_TOL, common.GRAD_TOL, full=False)
