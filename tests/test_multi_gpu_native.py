"""GPU: ptp_solve_batched_multi_* (one process, one host thread per device, NCCL gather to the root device) against
ptp_solve_batched_* on a single device, bit for bit. With one visible GPU only the host-rows path and the one-device
degenerate case run; `gpurun --gpus 2` exercises the NCCL transfer."""
import numpy as np
import pytest

from gproshan_b200 import api
from gproshan_b200 import meshgen as mg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    n_dev = api.device_count()
    assert n_dev > 0
    mesh = mg.icosphere(40, 2e-3, seed=3).astype(np.float32)
    srcs = mg.random_sources(99, 37, mesh.n_vertices, unique=True)    # 37: not a multiple of the device count
    with api.DeviceMesh(mesh, 0) as dm:
        want = dm.solve_batched(srcs)
    return n_dev, mesh, srcs, want


def test_multi_host_rows(setup):
    n_dev, mesh, srcs, want = setup
    meshes = [api.DeviceMesh(mesh, d) for d in range(n_dev)]
    try:
        got = api.solve_batched_multi(meshes, srcs)
        assert np.array_equal(got, want)
        st = meshes[0].last_stats
        assert st["ms_total"] > 0 and st["gpu_launches"] >= n_dev
    finally:
        for m in meshes:
            m.close()


def test_multi_device_rows_nccl_gather(setup):
    import torch
    n_dev, mesh, srcs, want = setup
    meshes = [api.DeviceMesh(mesh, d) for d in range(n_dev)]
    try:
        for chunks in (0, 1, 3):
            api.set_option("gather_chunks", chunks)
            rows = torch.zeros((srcs.size, mesh.n_vertices), dtype=torch.float32, device="cuda:0")
            api.solve_batched_multi(meshes, srcs, rows_device_ptr=rows.data_ptr())
            torch.cuda.synchronize()
            assert np.array_equal(rows.cpu().numpy(), want), f"gather_chunks={chunks}"
    finally:
        api.set_option("gather_chunks", 0)
        for m in meshes:
            m.close()


def test_multi_source_sets_and_errors(setup):
    n_dev, mesh, srcs, want = setup
    meshes = [api.DeviceMesh(mesh, d) for d in range(n_dev)]
    try:
        sets = [[int(srcs[0])], [int(srcs[1]), int(srcs[2])], [int(srcs[3])], [int(srcs[4]), int(srcs[5]), int(srcs[6])], [int(srcs[7])]]
        flat = np.concatenate(sets).astype(np.uint32)
        off = np.cumsum([0] + [len(s) for s in sets]).astype(np.uint64)
        got = api.solve_batched_multi(meshes, flat, off)
        with api.DeviceMesh(mesh, 0) as dm:
            ref = dm.solve_batched(flat, off)
        assert np.array_equal(got, ref)
        with pytest.raises(api.PtpError):
            api.solve_batched_multi(meshes + [meshes[0]], srcs)   # two handles on one device
    finally:
        for m in meshes:
            m.close()
