"""End-to-end parity of the CUDA path with the reference graph (BASELINE.json: 1e-3 mm per joint), through the C-ABI.

1. precision='strict' (float64 on CUDA cores, csrc/strict.cu) against the float64 CPU oracle on every BASELINE config:
   <= 1e-3 mm, plus layer by layer on config A; precision='strict_f16' against the oracle's 'half' mode (the float16
   graph with exact accumulation).  This pins the two strict evaluators to the oracle, which is itself pinned to the
   reference's code by tests/golden/ (tests/test_golden.py).
2. The tensor-core path -- the reference's default float16 export (src/options.py:73) -- is one legitimate evaluation
   of a graph whose result depends on summation order; its distance from the exact graph is gated STATISTICALLY over
   32 crops per config: mean and 99th percentile of |cuda - fp64| against those of |half - fp64|, the distance of the
   ideal float16 evaluation from the exact one.  The per-config table is written to gpurun_out/parity_table.json
   (copied to profiles/ by tools/collect_profiles.sh).
3. The rows SURVEY 8f marks "next" -- frozen-graph import, uint8 ingestion -- against the ORACLE (not against the CUDA
   path itself).
"""
import json
import os
import sys

import numpy as np
import pytest

from metro_pose3d_b200.joints import export_permutation, exported_joint_info, model_joint_info
from metro_pose3d_b200.spec import CONFIGS, NetSpec
from metro_pose3d_b200.weights import synth_images, synth_weights
from oracle.metro_oracle import OracleNet

pytestmark = pytest.mark.gpu

STRICT_TOL_MM = 1e-3          # BASELINE.json north_star: "within 1e-3 mm per joint"
# statistical gate of the float16 tensor-core path: its error distribution may not be wider than that of the ideal
# float16 evaluation by more than these factors (32 crops x J x 3 samples, fixed seeds, deterministic kernels)
MEAN_FACTOR, P99_FACTOR = 1.1, 1.1      # measured on the five configs: 0.94-0.99 / 0.92-1.00 (profiles/r2_parity_table.json)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _setup(cfg):
    arch, stride, ds, _, _ = CONFIGS[cfg]
    j = model_joint_info(ds).n_joints
    spec = NetSpec(arch, stride, j)
    return arch, stride, ds, spec, synth_weights(spec, 0), export_permutation(ds)


@pytest.mark.parametrize('prec,mode,tol', [('strict', 'fp64', 1e-11), ('strict_f16', 'half', 1e-11)])
def test_strict_precision_layer_by_layer(prec, mode, tol):
    """Every tensor of the strict evaluators against the oracle's trace (config A): float64 / the float16 graph with
    exact accumulation.  (Listed before the end-to-end tests so that a mismatch names its first layer.)"""
    import torch
    from metro_pose3d_b200.inference import MetroModel
    arch, stride, ds, spec, w, perm = _setup('A')
    img = synth_images(2, seed=1000)
    ora = OracleNet(spec, w, perm, mode)
    ora.trace = {}
    head = ora.forward_head(img)
    model = MetroModel(arch, stride, ds, weights=w, max_batch=2, precision=prec, keep_activations=True)
    model.infer(torch.from_numpy(img).cuda())
    torch.cuda.synchronize()
    report = []
    for name, t in list(ora.trace.items()) + [('head', head)]:
        if name == 'postnorm':
            continue
        got = model.debug_read(name).reshape(t.shape)
        report.append((name, float(np.abs(got - t).max() / max(np.abs(t).max(), 1e-30)), float((got != t).mean())))
    bad = [r for r in report if r[1] >= tol]
    assert not bad, f'first mismatching layers (name, max rel err, fraction of differing elements): {bad[:4]}'


@pytest.mark.parametrize('cfg,n', [('A', 2), ('B', 2), ('C', 2), ('D', 2), ('E', 1)])
def test_strict_precision_within_1e3_mm_of_the_fp64_oracle(cfg, n):
    import torch
    from metro_pose3d_b200.inference import MetroModel
    arch, stride, ds, spec, w, perm = _setup(cfg)
    img = synth_images(n, seed=1000)
    want = OracleNet(spec, w, perm, 'fp64')(img)
    model = MetroModel(arch, stride, ds, weights=w, max_batch=n, precision='strict')
    got = model.infer(torch.from_numpy(img).cuda()).cpu().numpy()
    err = float(np.abs(got - want).max())
    assert err <= STRICT_TOL_MM, f'config {cfg}: |strict - fp64 oracle| = {err:.3e} mm'
    host = model.infer_host(img)                       # the host-buffer call of a strict handle
    assert np.array_equal(host, got)
    model.close()
    # the float16 graph evaluated with exact accumulation, against the oracle's 'half' mode
    want16 = OracleNet(spec, w, perm, 'half')(img)
    m16 = MetroModel(arch, stride, ds, weights=w, max_batch=n, precision='strict_f16')
    got16 = m16.infer(torch.from_numpy(img).cuda()).cpu().numpy()
    err16 = float(np.abs(got16 - want16).max())
    assert err16 <= STRICT_TOL_MM, f'config {cfg}: |strict_f16 - half oracle| = {err16:.3e} mm'
    m16.close()


@pytest.mark.parametrize('cfg', ['A', 'B', 'C', 'D', 'E'])
def test_tensor_core_path_statistical_gate(cfg):
    """32 crops: the error of the float16 tensor-core path against the exact graph is distributed like that of the
    ideal float16 evaluation.  Both references are the strict evaluators pinned above (the float64 CPU oracle needs
    minutes for 32 crops of config E); two crops of each are re-checked against the CPU oracle here."""
    import torch
    from metro_pose3d_b200.inference import MetroModel
    arch, stride, ds, spec, w, perm = _setup(cfg)
    n = 32
    img = synth_images(n, seed=2000)
    x = torch.from_numpy(img).cuda()
    outs = {}
    for prec in ('strict', 'strict_f16', 'f16'):
        m = MetroModel(arch, stride, ds, weights=w, max_batch=n, precision=prec)
        outs[prec] = m.infer(x).cpu().numpy().astype(np.float64)
        m.close()
        torch.cuda.empty_cache()
    k = 2 if cfg != 'E' else 1
    assert np.abs(outs['strict'][:k] - OracleNet(spec, w, perm, 'fp64')(img[:k])).max() <= STRICT_TOL_MM
    e_cuda = np.abs(outs['f16'] - outs['strict']).ravel()
    e_half = np.abs(outs['strict_f16'] - outs['strict']).ravel()
    d_half = np.abs(outs['f16'] - outs['strict_f16']).ravel()
    row = {'config': cfg, 'crops': n, 'samples': int(e_cuda.size),
           'cuda_vs_fp64_mm': {'mean': e_cuda.mean(), 'p99': np.percentile(e_cuda, 99), 'max': e_cuda.max()},
           'half_vs_fp64_mm': {'mean': e_half.mean(), 'p99': np.percentile(e_half, 99), 'max': e_half.max()},
           'cuda_vs_half_mm': {'mean': d_half.mean(), 'p99': np.percentile(d_half, 99), 'max': d_half.max()},
           'pose_abs_max_mm': float(np.abs(outs['strict']).max())}
    print('parity', json.dumps(row, default=float))
    out_dir = os.path.join(ROOT, 'gpurun_out')
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, 'parity_table.json')
    table = json.load(open(path)) if os.path.exists(path) else {}
    table[cfg] = row
    json.dump(table, open(path, 'w'), indent=1, default=float)
    assert e_cuda.mean() <= MEAN_FACTOR * e_half.mean(), row
    assert np.percentile(e_cuda, 99) <= P99_FACTOR * np.percentile(e_half, 99), row


def test_frozen_graph_import_against_the_oracle(tmp_path):
    """a frozen GraphDef with the reference's node names -> TensorFlow-free importer -> CUDA path,
    against the ORACLE evaluating the weights the graph was written from: strict precision at 1e-3 mm; the joint
    tables come back through the C-ABI (metro_get_joint_info)."""
    import torch
    sys.path.insert(0, os.path.dirname(__file__))
    from pb_writer import frozen_graph
    from metro_pose3d_b200.inference import MetroModel, estimate_pose
    spec = NetSpec('resnet_v2_50', 32, 17)
    w = synth_weights(spec, 11)
    perm = export_permutation('h36m')
    ji = exported_joint_info('h36m')
    path = tmp_path / 'model.pb'
    path.write_bytes(frozen_graph(spec, w, perm, list(ji.names), np.asarray(ji.edges)))
    img = synth_images(3, seed=5)
    want = OracleNet(spec, w, perm, 'fp64')(img)
    strict = MetroModel.from_frozen_graph(str(path), max_batch=3, precision='strict')
    got = strict.infer(torch.from_numpy(img).cuda()).cpu().numpy()
    assert np.abs(got - want).max() <= STRICT_TOL_MM
    assert strict.joint_names == list(ji.names) and np.array_equal(strict.joint_edges, np.asarray(ji.edges))
    strict.close()
    # the reference's call: estimate_pose(images, 'model.pb') (inference.py:31-43), tensor-core path
    poses, edges, names = estimate_pose(torch.from_numpy(img).cuda(), str(path))
    half = OracleNet(spec, w, perm, 'half')(img)
    noise = np.abs(half - want).mean()
    assert np.abs(poses.cpu().numpy() - want).mean() <= 2.0 * noise + 0.05
    assert names == list(ji.names) and np.array_equal(edges, np.asarray(ji.edges)) and edges.dtype == np.int64


def test_uint8_ingestion_against_the_oracle():
    """uint8 crops through the strict path against the oracle fed np.float32(im) / 255
    (src/improc.py:56-61): 1e-3 mm."""
    import torch
    from metro_pose3d_b200.inference import MetroModel
    arch, stride, ds, spec, w, perm = _setup('A')
    u8 = np.random.default_rng(9).integers(0, 256, (2, 256, 256, 3), dtype=np.uint8)
    want = OracleNet(spec, w, perm, 'fp64')(u8.astype(np.float32) / np.float32(255))
    m = MetroModel(arch, stride, ds, weights=w, max_batch=2, precision='strict')
    got = m.infer(torch.from_numpy(u8).cuda()).cpu().numpy()
    assert np.abs(got - want).max() <= STRICT_TOL_MM
    assert np.array_equal(m.infer_host(u8), got)


def test_joint_tables_through_the_c_abi():
    """'joint_names' / 'joint_edges' (inference.py:36-38, main.py:128,140-141) via metro_get_joint_info."""
    from metro_pose3d_b200.inference import MetroModel
    for ds, arch, stride in (('h36m', 'resnet_v2_50', 32), ('coco19', 'resnet_v2_50', 32)):
        ji = exported_joint_info(ds)
        m = MetroModel(arch, stride, ds, max_batch=1)
        assert m.joint_names == list(ji.names)
        assert np.array_equal(m.joint_edges, np.asarray(ji.edges, dtype=np.int64))
        assert m.workspace_bytes(1) == m.workspace_bytes() > 0
        m.close()
    m = MetroModel('resnet_v2_50', 32, 'h36m', max_batch=8)
    assert m.workspace_bytes(2) < m.workspace_bytes(4) < m.workspace_bytes(8) == m.workspace_bytes()
    m.close()
    m = MetroModel('resnet_v2_50', 32, 'h36m', max_batch=1, permutation=[16, 0, 1])      # no tables given
    with pytest.raises(ValueError, match='joint tables'):
        m.joint_names


def test_two_devices_in_one_process():
    """Kernel attributes (dynamic shared memory opt-in) are per device: a second handle on another GPU of the same
    process must launch and agree bit for bit with the first."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs in one process (gpurun --gpus 2)')
    from metro_pose3d_b200.inference import MetroModel, SoftArgmax
    arch, stride, ds, spec, w, perm = _setup('A')
    img = synth_images(3, seed=4)
    outs = []
    for dev in (0, 1):
        m = MetroModel(arch, stride, ds, weights=w, max_batch=3, device=dev)
        with torch.cuda.device(dev):
            outs.append(m.infer(torch.from_numpy(img).cuda(dev)).cpu().numpy())
        m.close()
    assert np.array_equal(outs[0], outs[1])
