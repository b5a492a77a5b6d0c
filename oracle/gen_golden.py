#!/usr/bin/env python3
"""Generates tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN PYTHON (imported from
/root/reference/src) over oracle/tf_shim.py, a float64 numpy/torch-CPU stand-in for the TensorFlow
1.13 / slim entry points the inference graph touches.  TEST INFRASTRUCTURE.

    python oracle/gen_golden.py            # needs /root/reference; run in the build container only

The reference cannot run as shipped here (TensorFlow 1.13.1 is not installable), so this is the
closest available pin: every graph-level decision (block tables, centred stride, atrous rates,
padding, shortcut wiring, variable names and shapes, the heatmap reshape/transpose/softmax/decode
chain, metric scaling, root subtraction, export permutation, joint tables) is taken by the
reference's code; only the primitive op semantics come from the shim (see its header).

What main.export() does is replayed line by line in run_export() -- src/main.py:106-128 -- because
main.py itself imports the training / data stack (cv2, matplotlib, imageio, spacepy ...).
Modules executed from the reference: options, paths, util, tfu, tfu3d, data.datasets,
model.architectures, model.resnet_v2, model.resnet_utils, model.volumetric.
Stubbed (not on the inference path): init (replaced by a namespace filled by the real
options.get_parser()), util3d, data.datasets2d, model.bone_length_based_backproj's scipy use.

Inputs are the repo's seeded synthetic generators (metro_pose3d_b200/weights.py), so the fixtures
hold outputs only (plus the seeds); tests/test_golden.py regenerates the inputs and compares the
oracle (oracle/metro_oracle.py) and, on the GPU box, the CUDA path against them.
"""
from __future__ import annotations

import argparse
import ast
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('METRO_REFERENCE', '/root/reference')
sys.path.insert(0, ROOT)

from oracle import tf_shim  # noqa: E402


def _install_reference():
    """Put the reference's src/ on the path behind the shim and the three stubs."""
    if not os.path.isdir(os.path.join(REF, 'src')):
        raise SystemExit(f'{REF}/src not found: the golden generator only runs where the reference is mounted')
    tf = tf_shim.install()
    if os.path.join(REF, 'src') not in sys.path:
        sys.path.insert(0, os.path.join(REF, 'src'))
    import options                                      # reference: flag names and defaults
    flags = argparse.Namespace()
    options.get_parser().parse_args([], namespace=flags)
    init = types.ModuleType('init')                    # stands in for src/init.py (cv2/matplotlib/logging set-up)
    init.FLAGS = flags
    sys.modules['init'] = init
    sys.modules.setdefault('util3d', types.ModuleType('util3d'))
    if 'data.datasets2d' not in sys.modules:
        import data                                    # reference package (namespace)
        d2 = types.ModuleType('data.datasets2d')
        sys.modules['data.datasets2d'] = d2
        data.datasets2d = d2
    import tfu
    import model.volumetric                            # pulls architectures, resnet_v2, resnet_utils, tfu3d, data.datasets
    return tf, flags, tfu


def _literal_from_function(path, func, names):
    """Evaluates the literal assignments `name = ...` inside `func` of a reference source file (the
    joint tables live inside dataset-building functions that need the datasets on disk)."""
    tree = ast.parse(open(path).read())
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == func:
            for st in node.body:
                if isinstance(st, ast.Assign) and isinstance(st.targets[0], ast.Name) and st.targets[0].id in names:
                    out[st.targets[0].id] = eval(compile(ast.Expression(st.value), path, 'eval'), {})
    return out


def reference_joint_info(dataset):
    """JointInfo built by the reference's class from the reference's tables."""
    import data.datasets as ps3d
    src = os.path.join(REF, 'src', 'data')
    if dataset == 'h36m':
        t = _literal_from_function(os.path.join(src, 'h36m.py'), 'make_h36m', ('joint_names', 'edges'))
        return ps3d.JointInfo(t['joint_names'], t['edges'])
    if dataset == 'merged':
        return ps3d.make_merged().joint_info
    raise ValueError(dataset)


def export_permutation_from_reference(dataset):
    """The literal lists in main.export() (src/main.py:119-125), read from the source text."""
    tree = ast.parse(open(os.path.join(REF, 'src', 'main.py')).read())
    perms = []
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == 'export':
            for sub in ast.walk(node):
                if isinstance(sub, ast.Assign) and getattr(sub.targets[0], 'id', None) == 'permutation':
                    perms.append(ast.literal_eval(sub.value))
    return dict(zip(('merged', 'h36m', 'mpi_inf_3dhp'), perms))[dataset]


class FixedJoints:
    """A joint_info with n joints for heads the public tables do not cover (BASELINE's 19-joint
    COCO/CMU head): only .n_joints is read by net_output_to_heatmap_and_coords (volumetric.py:227-229)."""

    def __init__(self, n):
        self.n_joints = n


def run_export(tf, flags, tfu, arch, stride, joint_info, permutation, images_nhwc, weights, proc_side=256,
               data_format='NCHW', centered_stride=True, trace=False):
    """main.export() up to the 'output' tensor (src/main.py:106-128), after init.initialize()'s
    graph-affecting lines (src/init.py:54-70)."""
    import attrdict
    import model.volumetric
    import tensorflow.contrib.slim as slim
    flags.architecture, flags.stride_test, flags.proc_side = arch, stride, proc_side
    flags.centered_stride, flags.data_format = centered_stride, data_format
    flags.dataset = 'merged'   # only consulted for bone-length scale recovery (volumetric.py:167-170); 'merged' needs no files
    # init.py:54-59 -- golden vectors are the float32-declared graph evaluated in float64
    tfu.set_data_format(flags.data_format)
    tfu.set_dtype(tf.float32)
    tfu.set_is_training(False)                       # main.py:188 (test()) / module default None: inference-mode BN
    prefix = f'MainPart/{arch}/'
    tf.shim.reset(weights, strip_prefix=prefix)
    if trace:
        tf.shim._S.trace = {}
    with slim.arg_scope(                             # init.py:63-70
            [slim.conv2d, slim.conv3d, slim.conv3d_transpose, slim.conv2d_transpose, slim.avg_pool2d,
             slim.separable_conv2d, slim.max_pool2d, slim.batch_norm, slim.spatial_softmax],
            data_format=tfu.data_format()):
        with slim.arg_scope([slim.avg_pool2d, slim.max_pool2d], padding='SAME'):
            t = attrdict.AttrDict()
            t.x = tf.convert_to_tensor(images_nhwc, dtype=tf.float32, name='input')   # placeholder feed, main.py:109-110
            t.x = tfu.nhwc_to_std(t.x)                                                 # main.py:111
            model.volumetric.build_inference_model(joint_info, tfu.TEST, t)            # main.py:113
            out = tf.gather(t.coords3d_pred_rootrel, permutation, axis=1, name='output')   # main.py:127
    res = {'output': out.value, 'softmaxed': t.softmaxed.value, 'created': tf.shim.created_variables()}
    if trace:
        res['trace'] = tf.shim._S.trace
        tf.shim._S.trace = None
    return res


def run_decode(tf, flags, tfu, head_nhwc, joint_info, permutation, stride, proc_side=256, data_format='NCHW',
               centered_stride=True):
    """net_output_to_heatmap_and_coords + heatmap_to_metric + root_relative + gather on a given head
    tensor (volumetric.py:227-235,303-306,203; tfu3d.py:23-25; main.py:127)."""
    import model.volumetric as V
    import tfu3d
    flags.stride_test, flags.proc_side, flags.centered_stride = stride, proc_side, centered_stride
    tfu.set_data_format(data_format)
    net_output = tfu.nhwc_to_std(tf.convert_to_tensor(head_nhwc, dtype=tf.float32))
    softmaxed, coords3d = V.net_output_to_heatmap_and_coords(net_output, joint_info)
    metric = V.heatmap_to_metric(coords3d, tfu.TEST)
    out = tf.gather(tfu3d.root_relative(metric), permutation, axis=1)
    return out.value, coords3d.value


def run_to_orig_cam(tf, poses, rot, joint_info):
    """volumetric.to_orig_cam (src/model/volumetric.py:277-282) on given poses / rotations."""
    import model.volumetric as V
    return V.to_orig_cam(tf.convert_to_tensor(poses), tf.convert_to_tensor(rot), joint_info).value


def run_true_root_depth(tf, flags, tfu, coords01, inv_intrinsics, root_z, stride, proc_side=256, centered_stride=True):
    """The 'true-root-depth' branch of build_inference_model (src/model/volumetric.py:190-198) followed by root_relative
    (:203): image coordinates -> homogeneous -> inverse intrinsics -> back_project with the given root depth."""
    import model.volumetric as V
    import tfu3d
    flags.stride_test, flags.proc_side, flags.centered_stride = stride, proc_side, centered_stride
    coords3d = tf.convert_to_tensor(coords01)
    inv_k = tf.convert_to_tensor(inv_intrinsics)
    coords2d_pred = coords3d[..., :2]
    im_pred2d = V.heatmap_to_image(coords2d_pred, tfu.TEST)
    im_pred2d_homog = V.to_homogeneous_coords(im_pred2d)
    camcoords2d_homog = V.matmul_joint_coords(inv_k, im_pred2d_homog)
    delta_z_pred = (coords3d[..., 2] - coords3d[:, -1:, 2]) * flags.box_size_mm
    pred = V.back_project(camcoords2d_homog, delta_z_pred, tf.convert_to_tensor(root_z))
    return pred.value, tfu3d.root_relative(pred).value


def run_heatmap_pred_z(tf, flags, tfu, head_nhwc, joint_info):
    """t.heatmap_pred_z = reduce_sum(softmaxed, axis=[2, 3]) (src/model/volumetric.py:165): the depth marginal [N, J, D]."""
    import model.volumetric as V
    tfu.set_data_format('NCHW')
    net_output = tfu.nhwc_to_std(tf.convert_to_tensor(head_nhwc, dtype=tf.float32))
    softmaxed, _ = V.net_output_to_heatmap_and_coords(net_output, joint_info)
    return tf.reduce_sum(softmaxed, axis=[2, 3]).value


def synth_inv_intrinsics(n, seed, proc_side=256):
    """Inverse intrinsic matrices of crop cameras: focal length of a few hundred pixels, principal point near the centre."""
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        f = rng.uniform(250.0, 900.0)
        k = np.array([[f, 0.0, proc_side / 2 + rng.uniform(-4, 4)], [0.0, f * rng.uniform(0.98, 1.02), proc_side / 2 + rng.uniform(-4, 4)],
                      [0.0, 0.0, 1.0]])
        out.append(np.linalg.inv(k))
    return np.stack(out)


def synth_frame(h, w, seed):
    """A uint8 RGB frame: smooth gradients + blocks + noise, so that interpolation, borders and rounding all matter."""
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[:h, :w]
    base = np.stack([(xx * 255 // max(w - 1, 1)), (yy * 255 // max(h - 1, 1)), ((xx // 16 + yy // 16) % 2) * 200], -1)
    return np.clip(base + rng.randint(-40, 41, (h, w, 3)), 0, 255).astype(np.uint8)


def synth_crop_cameras(n, h, w, seed, side):
    """(intrinsics_old, R_old, intrinsics_new, R_new) per crop: a frame camera and a crop camera that was turned towards a
    point of the frame, zoomed and rolled -- what load_and_transform3d builds (data_loading.py:43-58), some crops reaching
    beyond the frame so that the constant border takes part."""
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        f = rng.uniform(0.8, 1.4) * w
        k_old = np.array([[f, 0, w / 2 + rng.uniform(-8, 8)], [0, f, h / 2 + rng.uniform(-8, 8)], [0, 0, 1.0]])
        q, r = np.linalg.qr(rng.randn(3, 3))
        r_old = q * np.sign(np.diag(r))
        if np.linalg.det(r_old) < 0:
            r_old[:, 0] = -r_old[:, 0]
        ang = rng.uniform(-0.35, 0.35, 3)
        cx, sx = np.cos(ang[0]), np.sin(ang[0]); cy, sy = np.cos(ang[1]), np.sin(ang[1]); cz, sz = np.cos(ang[2]), np.sin(ang[2])
        turn = (np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]) @
                np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]))
        fz = f * rng.uniform(0.5, 2.5) * side / w
        k_new = np.array([[fz, 0, (side - 1) / 2], [0, fz, (side - 1) / 2], [0, 0, 1.0]])
        out.append((k_old, r_old, k_new, turn @ r_old))
    return out


def run_reproject_image_fast(frame, cams, side):
    """The reference's own cameralib.reproject_image_fast (src/cameralib.py:406-429), cv2.remap included; cameralib's
    unused imports that are absent here (transforms3d) are stubbed."""
    sys.modules.setdefault('transforms3d', types.ModuleType('transforms3d'))
    import cameralib
    res = []
    for k_old, r_old, k_new, r_new in cams:
        old = types.SimpleNamespace(intrinsic_matrix=k_old, R=r_old)
        new = types.SimpleNamespace(intrinsic_matrix=k_new, R=r_new)
        res.append(cameralib.reproject_image_fast(frame, old, new, (side, side)))
    return np.stack(res)


def synth_rotations(n, seed):
    """n 3x3 matrices: proper rotations, every second one composed with a horizontal flip (det < 0), which is
    what `rot_to_orig_cam = ex.camera.R @ cam.R.T` holds after cam.horizontal_flip() (data_loading.py:80-83,110)."""
    rng = np.random.RandomState(seed)
    out = []
    for i in range(n):
        q, r = np.linalg.qr(rng.randn(3, 3))
        q = q * np.sign(np.diag(r))
        if np.linalg.det(q) < 0:
            q[:, 0] = -q[:, 0]
        if i % 2 == 1:
            q = q @ np.diag([-1.0, 1.0, 1.0])
        out.append(q)
    return np.stack(out)


def main(only=None):
    from metro_pose3d_b200.spec import NetSpec
    from metro_pose3d_b200.weights import synth_head, synth_images, synth_weights
    tf, flags, tfu = _install_reference()
    out_dir = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(out_dir, exist_ok=True)

    # ---- joint tables and permutations ------------------------------------------------------------
    tables = {}
    for ds in ('h36m', 'merged'):
        ji = reference_joint_info(ds)
        perm = export_permutation_from_reference(ds)
        tables[f'{ds}_model_names'] = np.array(ji.names)
        tables[f'{ds}_model_edges'] = np.array(ji.stick_figure_edges, dtype=np.int64)
        tables[f'{ds}_permutation'] = np.array(perm, dtype=np.int64)
        tables[f'{ds}_model_mirror'] = np.array(ji.mirror_mapping, dtype=np.int64)
        if ds == 'h36m':       # permute_joints needs every edge endpoint selected (true for h36m)
            pj = ji.permute_joints(perm)
            tables[f'{ds}_export_names'] = np.array(pj.names)
            tables[f'{ds}_export_edges'] = np.array(pj.stick_figure_edges, dtype=np.int64)
            tables[f'{ds}_export_mirror'] = np.array(pj.mirror_mapping, dtype=np.int64)
    np.savez_compressed(os.path.join(out_dir, 'joints.npz'), **tables)

    # ---- post-path: to_orig_cam on seeded poses and rotations -----------------------
    post = {}
    for ds in ('h36m', 'merged'):
        ji = reference_joint_info(ds)
        n = 6
        rng = np.random.RandomState(7 + ji.n_joints)
        poses = rng.randn(n, ji.n_joints, 3) * 400.0
        rot = synth_rotations(n, 11 + ji.n_joints)
        post[f'{ds}_orig_cam'] = run_to_orig_cam(tf, poses, rot, ji)
        post[f'{ds}_meta'] = np.array([n, ji.n_joints, 7 + ji.n_joints, 11 + ji.n_joints])
        # absolute-scale variant 'true-root-depth' (volumetric.py:190-198): seeded heatmap coordinates, cameras, depths
        rng = np.random.RandomState(23 + ji.n_joints)
        coords01 = rng.rand(n, ji.n_joints, 3)
        inv_k = synth_inv_intrinsics(n, 29 + ji.n_joints)
        root_z = rng.uniform(2000.0, 6000.0, n)
        for stride in (4, 16, 32):
            absolute, rootrel = run_true_root_depth(tf, flags, tfu, coords01, inv_k, root_z, stride)
            post[f'{ds}_trd_abs_s{stride}'], post[f'{ds}_trd_rel_s{stride}'] = absolute, rootrel
    # depth marginal of the softmaxed heatmap, t.heatmap_pred_z (volumetric.py:165)
    for name, side, j in (('A', 8, 17), ('C', 32, 19)):
        head = synth_head(2, side, j, seed=300 + side + j)
        post[f'pred_z_{name}'] = run_heatmap_pred_z(tf, flags, tfu, head, FixedJoints(j))
        post[f'pred_z_{name}_meta'] = np.array([2, side, j, 300 + side + j])
    np.savez_compressed(os.path.join(out_dir, 'post.npz'), **post)

    # ---- pre-path: crop extraction, the reference's reproject_image_fast --------------------------
    from oracle.crop_oracle import crop_homography, reproject_image_fast_ref
    crops = {}
    for name, h, w, side, n, seed in (('vga', 480, 640, 96, 5, 41), ('tall', 700, 380, 64, 4, 43)):
        frame = synth_frame(h, w, seed)
        cams = synth_crop_cameras(n, h, w, seed + 1, side)
        ref = run_reproject_image_fast(frame, cams, side)
        hs = np.stack([crop_homography(*c) for c in cams])
        mine = np.stack([reproject_image_fast_ref(frame, hm, side, side) for hm in hs])
        crops[f'{name}_crops'] = ref
        crops[f'{name}_homographies'] = hs
        crops[f'{name}_meta'] = np.array([h, w, side, n, seed])
        # pixels where the reference's BLAS-ordered float32 coordinates round to another 1/32-pixel cell than the fixed
        # evaluation order of oracle/crop_oracle.py (documented there); recorded so that the GPU test can hold the rest exact
        crops[f'{name}_blas_mismatch'] = np.argwhere(np.any(ref != mine, axis=-1))
        print(f'crops {name}: {ref.shape}, border pixels {(ref.sum(-1) == 0).mean():.2%}, '
              f'pixels differing from the fixed-order oracle: {len(crops[f"{name}_blas_mismatch"])} of {ref[..., 0].size}', flush=True)
    np.savez_compressed(os.path.join(out_dir, 'crops.npz'), **crops)
    if only == 'post':
        return

    # ---- decode only: every stride / joint-set / layout the configs use ------------------------------
    dec = {}
    h36m = reference_joint_info('h36m')
    p_h36m = export_permutation_from_reference('h36m')
    p_merged = export_permutation_from_reference('merged')
    cases = [('A', 8, 32, 17, h36m, p_h36m), ('B', 16, 16, 17, h36m, p_h36m), ('C', 32, 8, 19, FixedJoints(19), p_merged),
             ('D', 16, 16, 19, FixedJoints(19), p_merged), ('E', 64, 4, 19, FixedJoints(19), p_merged),
             ('M', 8, 32, 53, reference_joint_info('merged'), p_merged)]
    for name, side, stride, j, ji, perm in cases:
        n = 3 if side < 64 else 2
        head = synth_head(n, side, j, seed=100 + side + j)
        for fmt in ('NCHW', 'NHWC'):
            poses, c01 = run_decode(tf, flags, tfu, head, ji, perm, stride, data_format=fmt)
            if fmt == 'NCHW':
                dec[f'{name}_poses'], dec[f'{name}_coords01'] = poses, c01
            else:
                assert np.array_equal(poses, dec[f'{name}_poses']), 'decode must not depend on data_format'
        dec[f'{name}_meta'] = np.array([n, side, stride, j, 100 + side + j])
        # non-centred variant (FLAGS.centered_stride False) differs only by an offset that cancels root-relative
        p2, _ = run_decode(tf, flags, tfu, head, ji, perm, stride, centered_stride=False)
        assert np.allclose(p2, poses, atol=1e-9)
    np.savez_compressed(os.path.join(out_dir, 'decode.npz'), **dec)

    # ---- whole graph: images -> poses, with per-layer statistics -------------------------------------
    nets = [('rn50_s32', 'resnet_v2_50', 32, 256, 2), ('rn50_s16', 'resnet_v2_50', 16, 256, 1),
            ('rn50_s8', 'resnet_v2_50', 8, 128, 1), ('rn50_s4', 'resnet_v2_50', 4, 64, 1),
            ('rn101_s16', 'resnet_v2_101', 16, 128, 1), ('rn101_s4', 'resnet_v2_101', 4, 64, 1),
            ('rn50_s16_nocenter', 'resnet_v2_50', 16, 128, 1), ('rn101_s32', 'resnet_v2_101', 32, 128, 1)]
    g = {}
    for name, arch, stride, side, n in nets:
        centered = 'nocenter' not in name
        spec = NetSpec(arch, stride, 17, centered_stride=centered, proc_side=side)
        w = synth_weights(spec, seed=0)
        img = synth_images(n, seed=1000, side=side)
        res = run_export(tf, flags, tfu, arch, stride, h36m, p_h36m, img, w, proc_side=side, centered_stride=centered,
                         trace=True)
        # the NHWC build of the same graph must agree (src/options.py:92 makes the layout a flag)
        res2 = run_export(tf, flags, tfu, arch, stride, h36m, p_h36m, img, w, proc_side=side, data_format='NHWC',
                          centered_stride=centered)
        assert np.allclose(res['output'], res2['output'], rtol=0, atol=1e-7), name
        g[f'{name}_poses'] = res['output']
        g[f'{name}_meta'] = np.array([n, side, stride, 17, int(centered), 0, 1000])
        g[f'{name}_vars'] = np.array([k for k, _ in res['created']])
        g[f'{name}_var_shapes'] = np.array([','.join(map(str, s)) for _, s in res['created']])
        keys = sorted(res['trace'])
        g[f'{name}_layers'] = np.array(keys)
        g[f'{name}_layer_shapes'] = np.array([','.join(map(str, res['trace'][k].shape)) for k in keys])
        g[f'{name}_layer_rms'] = np.array([float(np.sqrt(np.mean(res['trace'][k] ** 2))) for k in keys])
        g[f'{name}_layer_sum'] = np.array([float(np.sum(res['trace'][k])) for k in keys])
        # a thin slice of every layer (first crop, first row, first 8 channels) pins position-dependent errors
        g[f'{name}_layer_probe'] = np.array([np.pad(res['trace'][k][0, 0, :, :8].ravel()[:64], (0, max(0, 64 - res['trace'][k][0, 0, :, :8].size)))
                                             for k in keys])
        print(f'{name}: {len(res["created"])} variables, {len(keys)} traced tensors, poses {res["output"].shape}', flush=True)
    np.savez_compressed(os.path.join(out_dir, 'graph.npz'), **g)
    print('wrote', out_dir)


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else None)     # `gen_golden.py post`: joint tables and post-path vectors only
