"""The C-ABI library loads on a CPU-only box and exports every symbol include/ddk.h declares; the two host builds of
device routines (Kabsch alignment, axis-angle) agree with numpy / the oracle.  No GPU compute here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from disco_diffdock_b200 import build, engine
from oracle import restate

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    build.build()
    return engine.load_library()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, 'include', 'ddk.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    names = set(re.findall(r'\b(ddk_\w+)\s*\(', hdr))
    assert names == set(engine.EXPORTS), names ^ set(engine.EXPORTS)
    for n in names:
        assert hasattr(lib, n), n
    assert lib.ddk_abi_version() == 1


def test_create_fails_loudly_without_gpu(lib):
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from tests import helpers
    m, sd, cfg = helpers.make_model(0)
    with pytest.raises(RuntimeError):
        m.engine('cuda')
    from disco_diffdock_b200 import data as ddata
    _, lst = helpers.make_pose_batch(3, 8, 12, 1)
    batch = ddata.Batch.from_data_list(lst)
    restate.set_time(batch, 0.5, 0.5, 0.5, 1)
    with pytest.raises(RuntimeError):
        m(batch)


def kabsch_numpy(A, B):
    ca, cb = A.mean(0), B.mean(0)
    H = (A - ca).T @ (B - cb)
    U, S, Vt = np.linalg.svd(H)
    R = Vt.T @ U.T
    if np.linalg.det(R) < 0:
        R = Vt.T @ np.diag([1., 1., -1.]) @ U.T
    return R, -R @ ca + cb


@pytest.mark.parametrize('case', ['generic', 'near_identity', 'reflection', 'planar'])
def test_host_kabsch_matches_svd(lib, case):
    rng = np.random.default_rng(1)
    A = rng.normal(size=(30, 3)) * 4
    if case == 'planar':
        A[:, 2] = 0
    Q = np.linalg.qr(rng.normal(size=(3, 3)))[0]
    if np.linalg.det(Q) < 0:
        Q[:, 0] *= -1
    if case == 'near_identity':
        B = A + rng.normal(size=A.shape) * 0.05
    elif case == 'reflection':
        B = -A + rng.normal(size=A.shape) * 0.01
    else:
        B = A @ Q.T + rng.normal(size=A.shape) * 0.3 + np.array([3., -2., 7.])
    A32, B32 = np.ascontiguousarray(A, np.float32), np.ascontiguousarray(B, np.float32)
    R, t = np.zeros(9, np.float32), np.zeros(3, np.float32)
    rc = lib.ddk_host_kabsch(A32.ctypes.data, B32.ctypes.data, 30, R.ctypes.data, t.ctypes.data)
    assert rc == 0
    Rn, tn = kabsch_numpy(A32.astype(np.float64), B32.astype(np.float64))
    if case in ('reflection', 'planar'):      # degenerate optimum: compare the residual instead of the matrix
        res = np.linalg.norm(A32 @ R.reshape(3, 3).T + t - B32)
        resn = np.linalg.norm(A32 @ Rn.T + tn - B32)
        assert res <= resn * (1 + 1e-4) + 1e-4
        assert abs(np.linalg.det(R.reshape(3, 3).astype(np.float64)) - 1) < 1e-5
    else:
        assert np.abs(R.reshape(3, 3) - Rn).max() < 2e-6
        assert np.abs(t - tn).max() < 2e-5


def test_host_axis_angle_matches_oracle(lib):
    g = torch.Generator().manual_seed(0)
    aa = torch.randn(64, 3, generator=g) * torch.tensor([1e-8, 0.3, 3.0]).repeat(64, 1)[:, :1].T.reshape(-1)[:64, None]
    aa[0] = 0
    ref = restate.axis_angle_to_matrix(aa)
    for i in range(64):
        v = np.ascontiguousarray(aa[i].numpy(), np.float32)
        R = np.zeros(9, np.float32)
        assert lib.ddk_host_axis_angle_to_matrix(v.ctypes.data, R.ctypes.data) == 0
        assert np.abs(R.reshape(3, 3) - ref[i].numpy()).max() < 1e-6


def test_lane_tables_cover_every_basis_row(lib):
    """k_conv_fused: every FasterTensorProduct basis row of every level is owned by exactly one (slot, lane)."""
    lib.ddk_host_lane_tables_check.restype = ctypes.c_int
    assert lib.ddk_host_lane_tables_check() == 0


def test_tensor_core_row_table_matches_basis(lib):
    """k_acc_tc: the (type, source column, harmonic) row table, evaluated on the host, reproduces the FasterTensorProduct basis
    in kernel order at every level; every row appears once and the table is sorted by row type."""
    from tests.test_weights_packing import raw_basis
    lib.ddk_host_tc_rows_eval.restype = ctypes.c_int
    lib.ddk_host_tc_rows_eval.argtypes = [ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    g = torch.Generator().manual_seed(0)
    x, sh = torch.randn(5, 84, generator=g), torch.randn(5, 4, generator=g)
    for lv, U in enumerate((96, 138, 180, 276)):
        want = raw_basis(x, sh, lv).numpy()
        assert want.shape[1] == U
        for i in range(5):
            xi = np.ascontiguousarray(x[i].numpy(), np.float32)
            si = np.ascontiguousarray(sh[i].numpy(), np.float32)
            out = np.zeros(U, np.float32)
            assert lib.ddk_host_tc_rows_eval(lv, xi.ctypes.data, si.ctypes.data, out.ctypes.data) == U
            assert np.abs(out - want[i]).max() < 1e-5


def test_tf32_split_is_exact(lib):
    """k_acc_tc's operand split: hi on the TF32 grid, hi + lo == a exactly, |lo| <= half a TF32 ulp of a, so that the three-pass
    product hi*hi + hi*lo + lo*hi drops only terms of order 2^-21 (tools/microbench/umma_tf32x3.cu measures 4-6e-7)."""
    lib.ddk_host_tc_split.restype = ctypes.c_int
    lib.ddk_host_tc_split.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]
    rng = np.random.default_rng(0)
    a = np.concatenate([rng.standard_normal(4096) * 10.0 ** rng.integers(-6, 6, 4096), [0.0, -0.0, 1.0, -1.0, 1.0 + 2.0 ** -11,
                        1.0 + 2.0 ** -12, 3.4e38 * 0.5, 1e-30, -7.25]]).astype(np.float32)
    hi, lo = np.zeros(a.size, np.uint32), np.zeros(a.size, np.uint32)
    assert lib.ddk_host_tc_split(a.ctypes.data, a.size, hi.ctypes.data, lo.ctypes.data) == 0
    assert not (hi & 0x1fff).any()
    h, l = hi.view(np.float32), lo.view(np.float32)
    assert np.array_equal(h + l, a)                                        # exact in fp32
    assert (np.abs(l.astype(np.float64)) <= np.abs(a.astype(np.float64)) * 2.0 ** -11 * (1 + 2.0 ** -10) + 1e-45).all()
    assert np.array_equal(np.signbit(h[np.abs(a) > 0]), np.signbit(a[np.abs(a) > 0]))


def test_tensor_core_role_tables(lib):
    """k_conv_tcr: at every basis level each (basis row, hidden unit) and (basis row, bias) pair belongs to exactly one role, the
    resident weight slices reproduce the packed second-layer weights, the shapes fit tensor memory / shared memory, and the partial
    records of the contraction warps summed through the finalize table equal the direct contraction (emulated on the host)."""
    lib.ddk_host_tcr_roles_check.restype = ctypes.c_int
    assert lib.ddk_host_tcr_roles_check() == 0


def test_tf32_split_rounded_lo(lib):
    """k_conv_tcr's operand split: hi and lo both on the TF32 grid (nothing is left to the tensor core's truncation),
    |a - hi - lo| <= 2^-22 |a|."""
    lib.ddk_host_tc_split_rn.restype = ctypes.c_int
    lib.ddk_host_tc_split_rn.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]
    rng = np.random.default_rng(1)
    a = np.concatenate([rng.standard_normal(4096) * 10.0 ** rng.integers(-6, 6, 4096), [0.0, -0.0, 1.0, -1.0, 1.0 + 2.0 ** -11,
                        1.0 + 2.0 ** -12, 1.0 + 2.0 ** -23, 1e-30, -7.25]]).astype(np.float32)
    hi, lo = np.zeros(a.size, np.uint32), np.zeros(a.size, np.uint32)
    assert lib.ddk_host_tc_split_rn(a.ctypes.data, a.size, hi.ctypes.data, lo.ctypes.data) == 0
    assert not (hi & 0x1fff).any() and not (lo & 0x1fff).any()
    h, l = hi.view(np.float32).astype(np.float64), lo.view(np.float32).astype(np.float64)
    assert (np.abs(a.astype(np.float64) - h - l) <= np.abs(a.astype(np.float64)) * 2.0 ** -22 + 1e-45).all()
