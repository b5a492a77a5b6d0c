"""GPU parity of the whole path (crops -> poses) through the C-ABI against the CPU oracle."""
import numpy as np
import pytest

from metro_pose3d_b200.joints import export_permutation
from metro_pose3d_b200.spec import NetSpec
from metro_pose3d_b200.weights import synth_weights, synth_images
from oracle.metro_oracle import OracleNet

pytestmark = pytest.mark.gpu

# End-to-end tolerances (BASELINE.json: 1e-3 mm per joint against the reference graph).
#   * precision='strict' (float64 on CUDA cores) meets 1e-3 mm against the float64 oracle on every config:
#     tests/test_parity_gpu.py.
#   * the tensor-core path is the reference's DEFAULT float16 graph (src/options.py:73), whose output is only
#     defined up to the summation order of its fp16 convolutions: the oracle that rounds to fp16 at the same
#     storage points ('half' mode) and the CUDA path differ wherever fp32-vs-exact accumulation flips an
#     individual fp16 rounding, and those flips random-walk through ~50 layers.  Here it is checked layer by layer
#     (relative L2 per tensor) and, for the decode, at 1e-3 mm on identical logits; its end-to-end distance from the
#     float64 graph is gated statistically over 32 crops per config in tests/test_parity_gpu.py.
LAYER_REL = 2e-3
DECODE_TOL_MM = 1e-3


def _rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize('arch,stride,ds,n', [('resnet_v2_50', 32, 'h36m', 2), ('resnet_v2_50', 16, 'h36m', 2),
                                             ('resnet_v2_101', 16, 'coco19', 1), ('resnet_v2_50', 8, 'coco19', 1),
                                             ('resnet_v2_50', 4, 'coco19', 1), ('resnet_v2_101', 4, 'coco19', 1)])
def test_layerwise_and_end_to_end(arch, stride, ds, n):
    import torch
    from metro_pose3d_b200.inference import MetroModel, estimate_pose
    from metro_pose3d_b200.joints import model_joint_info
    perm = export_permutation(ds)
    j = model_joint_info(ds).n_joints
    spec = NetSpec(arch, stride, j)
    w = synth_weights(spec, 0)
    img = synth_images(n, seed=1000)
    model = MetroModel(arch, stride, ds, weights=w, max_batch=n, keep_activations=True)
    poses, edges, names = estimate_pose(torch.from_numpy(img).cuda(), model)
    torch.cuda.synchronize()
    poses = poses.cpu().numpy()
    ora = OracleNet(spec, w, perm, 'half')
    ora.trace = {}
    head = ora.forward_head(img)
    ref = ora.decode(head)
    report = []
    for name, t in ora.trace.items():
        if name == 'postnorm':
            continue
        got = model.debug_read(name).reshape(t.shape)
        report.append((name, _rel(got, t)))
    got_head = model.debug_read('head').reshape(head.shape)
    report.append(('head', _rel(got_head, head)))
    worst = max(report, key=lambda r: r[1])
    # rounding flips random-walk with depth: 105 convolutions instead of 54 -> sqrt(2) wider band for ResNet-101
    layer_rel = LAYER_REL * (1.5 if arch.endswith('101') else 1.0)
    assert worst[1] < layer_rel, f'worst layer {worst}; first bad: {[r for r in report if r[1] >= layer_rel][:3]}'
    # strict: the decode of the path's own head tensor
    strict = np.abs(poses - ora.decode(got_head)).max()
    assert strict < DECODE_TOL_MM, f'decode on identical logits: max |err| = {strict:.3e} mm'
    # end to end: a coarse sanity bound here (1-2 crops); the statistical gate is tests/test_parity_gpu.py
    err_half = np.abs(poses - ref).max()
    assert err_half < 10.0, f'end-to-end: |cuda-half| {err_half:.3f} mm'
    assert poses.shape == (n, len(perm), 3) and len(names) == len(perm)
    if ds == 'h36m':
        assert np.all(poses[:, 0] == 0) and edges.shape == (16, 2) and names[0] == 'pelv'


def test_host_buffer_call_and_batch_invariance():
    """metro_infer_host (numpy in / numpy out) equals the device-buffer call; a crop's result does not
    depend on its batch position or on the batch it is sharded into (the multi-GPU parity property)."""
    import torch
    from metro_pose3d_b200.inference import MetroModel
    spec = NetSpec('resnet_v2_50', 32, 17)
    w = synth_weights(spec, 0)
    img = synth_images(5, seed=1001)
    model = MetroModel('resnet_v2_50', 32, 'h36m', weights=w, max_batch=8)
    a = model.infer(torch.from_numpy(img).cuda()).cpu().numpy()
    b = model.infer_host(img)
    assert np.array_equal(a, b)
    c = model.infer_host(img[[4, 2, 0, 1, 3]])
    assert np.array_equal(c, a[[4, 2, 0, 1, 3]])
    d = np.concatenate([model.infer_host(img[:3]), model.infer_host(img[3:])])
    assert np.array_equal(d, a)
    with pytest.raises(ValueError):
        model.infer_host(img[:, :128])
    with pytest.raises(ValueError):
        model.infer_host(img, out=np.zeros((4, 17, 3), np.float32))   # undersized output buffer
    # the graph's placeholder is [None,256,256,3]: batches above max_batch run in pieces of max_batch crops
    big = synth_images(19, seed=1002)
    e = model.infer_host(big)
    f = np.concatenate([model.infer_host(big[:8]), model.infer_host(big[8:16]), model.infer_host(big[16:])])
    assert np.array_equal(e, f)
    assert np.array_equal(model.infer(torch.from_numpy(big).cuda()).cpu().numpy(), e)


def test_sliced_host_path_equals_device_path():
    """metro_infer_host runs the stem of the network slice by slice underneath the PCIe copies (64-crop
    slices, ragged last one) and the deep blocks on the whole batch: bit-identical to the device-buffer call."""
    import torch
    from metro_pose3d_b200.inference import MetroModel
    n = 160                                       # 64 + 64 + 32
    model = MetroModel('resnet_v2_50', 16, 'h36m', max_batch=n)
    img = torch.rand((n, 256, 256, 3), dtype=torch.float32)
    a = model.infer(img.cuda()).cpu().numpy()
    b = model.infer_host(img.numpy())
    assert np.array_equal(a, b)
    assert np.isfinite(a).all() and np.abs(a).max() > 1.0


def test_large_batch_matches_small_batches():
    """A crop's result must not depend on the batch it is part of, in particular not on where the persistent
    kernels' work lists wrap around the grid (148 CTAs / 74 CTA pairs): 160 crops in one call vs the same crops
    in calls of 8, bit for bit, and run-to-run determinism."""
    import torch
    from metro_pose3d_b200.inference import MetroModel
    n = 160
    model = MetroModel('resnet_v2_50', 16, 'h36m', max_batch=n)
    g = torch.Generator().manual_seed(5)
    img = torch.rand((n, 256, 256, 3), generator=g, dtype=torch.float32).cuda()
    full = model.infer(img).cpu().numpy()
    assert np.array_equal(full, model.infer(img).cpu().numpy())
    for lo in (0, 16, 72, 144, 152):
        part = model.infer(img[lo:lo + 8].contiguous()).cpu().numpy()
        assert np.array_equal(part, full[lo:lo + 8]), f'crops {lo}..{lo + 8} depend on their batch'


def test_in_kernel_preactivation_is_bit_identical():
    """Production handles let identity units of block1/2 read their raw input and apply the pre-activation
    inside conv1 (no stored pre-activation tensor); keep_activations handles store every tensor.  Same
    arithmetic and rounding points, so the poses must agree bit for bit."""
    import torch
    from metro_pose3d_b200.inference import MetroModel
    w = synth_weights(NetSpec('resnet_v2_50', 16, 17), 3)
    img = torch.from_numpy(synth_images(6, seed=77)).cuda()
    a = MetroModel('resnet_v2_50', 16, 'h36m', weights=w, max_batch=6).infer(img).cpu().numpy()
    b = MetroModel('resnet_v2_50', 16, 'h36m', weights=w, max_batch=6, keep_activations=True).infer(img).cpu().numpy()
    assert np.array_equal(a, b)


def test_frozen_graph_import_runs_the_same_model(tmp_path):
    """A frozen GraphDef written with the reference's node names (tests/pb_writer.py) and loaded through the
    TensorFlow-free importer gives the poses of the model built from the same weights; estimate_pose accepts the
    .pb path like the reference's inference.py:31-43."""
    import sys, os, torch
    sys.path.insert(0, os.path.dirname(__file__))
    from pb_writer import frozen_graph
    from metro_pose3d_b200.inference import MetroModel, estimate_pose
    from metro_pose3d_b200.joints import exported_joint_info
    spec = NetSpec('resnet_v2_50', 32, 17)
    w = synth_weights(spec, 11)
    ji = exported_joint_info('h36m')
    path = tmp_path / 'model.pb'
    path.write_bytes(frozen_graph(spec, w, export_permutation('h36m'), list(ji.names), np.asarray(ji.edges)))
    img = torch.from_numpy(synth_images(3, seed=5)).cuda()
    ref = MetroModel('resnet_v2_50', 32, 'h36m', weights=w, max_batch=3).infer(img).cpu().numpy()
    poses, edges, names = estimate_pose(img, str(path))
    assert np.array_equal(poses.cpu().numpy(), ref)
    assert names == list(ji.names) and np.array_equal(edges, np.asarray(ji.edges))


def test_uint8_ingestion_is_bit_identical_to_the_float_path():
    """uint8 crops (fused x * (1/255f), csrc/root_fused.cu) against float32 crops prepared exactly as the reference
    does (np.float32(im) / 255, src/improc.py:56-61): after the cast to float16 (architectures.py:29) the two agree
    for all 256 byte values, so conv1 and the poses must be bit-identical.  The first crop holds every byte value in
    every channel."""
    import torch
    from metro_pose3d_b200.inference import MetroModel
    model = MetroModel('resnet_v2_50', 32, 'h36m', max_batch=2, keep_activations=True)
    rng = np.random.default_rng(3)
    u8 = rng.integers(0, 256, (2, 256, 256, 3), dtype=np.uint8)
    u8[0] = (np.arange(256 * 256 * 3, dtype=np.int64).reshape(256, 256, 3) % 256).astype(np.uint8)
    assert all(len(np.unique(u8[0, :, :, c])) == 256 for c in range(3))
    f32 = u8.astype(np.float32) / np.float32(255)          # improc.normalize01
    a = model.infer(torch.from_numpy(u8).cuda()).cpu().numpy()
    conv1_a = model.debug_read('conv1').copy()
    b = model.infer(torch.from_numpy(f32).cuda()).cpu().numpy()
    conv1_b = model.debug_read('conv1')
    assert np.array_equal(conv1_a, conv1_b)
    assert np.array_equal(a, b)
    # and the cast itself, exhaustively: fp16(k * (1/255f)) == fp16(float32(k) / 255)
    k = np.arange(256, dtype=np.float32)
    assert np.array_equal((k * np.float32(1.0 / 255.0)).astype(np.float16), (k / np.float32(255)).astype(np.float16))


def test_uint8_host_path_equals_device_path():
    """metro_infer_host_u8 (uint8 crops over PCIe in slices) == metro_infer_u8 on the same crops, bit for bit, at a
    batch that is not a multiple of the slice sizes."""
    import torch
    from metro_pose3d_b200.inference import MetroModel
    n = 160
    model = MetroModel('resnet_v2_50', 16, 'h36m', max_batch=n)
    g = torch.Generator().manual_seed(5)
    u8 = torch.randint(0, 256, (n, 256, 256, 3), dtype=torch.uint8, generator=g)
    host = model.infer_host(u8.numpy())
    dev = model.infer(u8.cuda()).cpu().numpy()
    assert np.array_equal(host, dev)
    with pytest.raises(ValueError):
        model.infer_host(u8.numpy().astype(np.int32))


def test_cuda_graph_replay_is_bit_identical():
    """Small batches are replayed from a captured CUDA graph (metro_graph_stats) the second time the same buffers are
    seen; the results must equal the direct launches bit for bit, for new image CONTENTS in the same buffer too."""
    import torch
    from metro_pose3d_b200.inference import MetroModel
    w = synth_weights(NetSpec('resnet_v2_50', 32, 17), 0)
    model = MetroModel('resnet_v2_50', 32, 'h36m', weights=w, max_batch=4)
    imgs = [torch.from_numpy(synth_images(4, seed=50 + i)).cuda() for i in range(3)]
    direct, keep = [], []
    for x in imgs:                                       # distinct buffers every call (kept alive): never replayed
        keep.append((x.clone(), torch.empty((4, 17, 3), device='cuda')))
        direct.append(model.infer(keep[-1][0], out=keep[-1][1]).cpu().numpy())
    assert model.graph_stats() == (0, 0)
    buf = torch.empty_like(imgs[0])
    out = torch.empty((4, 17, 3), device='cuda')
    for rep in range(2):
        for i, x in enumerate(imgs):
            buf.copy_(x)
            got = model.infer(buf, out=out).cpu().numpy()
            assert np.array_equal(got, direct[i]), (rep, i)
    graphs, replays = model.graph_stats()
    assert graphs == 1 and replays == 5                  # first sight runs directly, the next five are replays
    # uint8 feed of the same batch size: its own graph
    u8 = torch.randint(0, 256, (4, 256, 256, 3), dtype=torch.uint8, device='cuda')
    a = model.infer(u8, out=out).cpu().numpy()
    b = model.infer(u8, out=out).cpu().numpy()
    assert np.array_equal(a, b) and model.graph_stats()[0] == 2
    del keep
