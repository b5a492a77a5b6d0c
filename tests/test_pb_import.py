"""The TensorFlow-free frozen-graph importer (metro_pose3d_b200/pb_import.py) against graphs written by
tests/pb_writer.py with the reference's node names, ops and attributes (src/main.py:106-160)."""
import numpy as np
import pytest

from metro_pose3d_b200.joints import export_permutation, exported_joint_info
from metro_pose3d_b200.pb_import import import_frozen_graph, parse_graph_def
from metro_pose3d_b200.spec import NetSpec
from metro_pose3d_b200.weights import pack_blob, synth_weights

from pb_writer import frozen_graph, node, attr_tensor


def _graph(arch, stride, ds, half=False, seed=0):
    ji = exported_joint_info(ds)
    perm = export_permutation(ds)
    spec = NetSpec(arch, stride, 17 if ds == 'h36m' else 19)
    w = synth_weights(spec, seed)
    data = frozen_graph(spec, w, perm, list(ji.names), np.asarray(ji.edges), half=half)
    return spec, w, perm, ji, data


@pytest.mark.parametrize('arch,stride,ds', [('resnet_v2_50', 32, 'h36m'), ('resnet_v2_50', 16, 'h36m'),
                                           ('resnet_v2_101', 8, 'coco19'), ('resnet_v2_50', 4, 'coco19')])
def test_round_trip_float32(arch, stride, ds):
    spec, w, perm, ji, data = _graph(arch, stride, ds)
    m = import_frozen_graph(data)
    assert (m.arch, m.stride, m.n_joints_model) == (arch, stride, spec.n_joints)
    assert m.permutation == list(perm)
    assert m.joint_names == list(ji.names)
    assert np.array_equal(m.joint_edges, np.asarray(ji.edges))
    assert list(m.weights) == list(w)
    assert np.array_equal(pack_blob(spec, m.weights), pack_blob(spec, w))


def test_half_constants_are_widened_exactly():
    spec, w, _, _, data = _graph('resnet_v2_50', 16, 'h36m', half=True, seed=3)
    m = import_frozen_graph(data)
    for name, a in w.items():
        assert m.weights[name].dtype == np.float32
        assert np.array_equal(m.weights[name], a.astype(np.float16).astype(np.float32)), name


def test_wire_format_details():
    """negative int64, scalar broadcast, unknown fields and attributes are handled / skipped."""
    g = node('a', 'Const', value=attr_tensor(np.asarray([-3, 7], np.int64), 'list'), dtype=b'\x30\x09') + \
        node('b', 'Identity', ['a:0'])
    nodes = parse_graph_def(g)
    assert list(nodes) == ['a', 'b'] and nodes['b'].inputs == ['a:0']
    assert nodes['a'].attr['value'].tolist() == [-3, 7]


def test_errors():
    spec, w, perm, ji, data = _graph('resnet_v2_50', 32, 'h36m')
    with pytest.raises(ValueError):
        import_frozen_graph(b'')                                     # no resnet scope
    broken = dict(w)
    broken['logits/weights'] = broken['logits/weights'][..., :100]    # head width not a multiple of depth
    bad = frozen_graph(spec, broken, perm, list(ji.names), np.asarray(ji.edges))
    with pytest.raises(ValueError):
        import_frozen_graph(bad)


@pytest.mark.parametrize('arch,stride', [('resnet_v2_50', 32), ('resnet_v2_101', 16)])
def test_folded_batch_norms_give_the_same_network(arch, stride):
    """A graph whose batch norms were built from primitive ops and folded (scale inside the filters, mul_1 / add_1
    for the pre-activations) imports as a blob that computes the same function: the oracle on the imported weights
    equals the oracle on the original ones, and every folded scale / shift equals the original's."""
    from oracle.metro_oracle import OracleNet
    from metro_pose3d_b200.weights import synth_images
    ji = exported_joint_info('h36m')
    perm = export_permutation('h36m')
    spec = NetSpec(arch, stride, 17)
    w = synth_weights(spec, 5)
    data = frozen_graph(spec, w, perm, list(ji.names), np.asarray(ji.edges), bn_form='folded')
    m = import_frozen_graph(data)
    assert list(m.weights) == list(w) and m.stride == stride and m.permutation == list(perm)
    eps = 1e-5
    for name in w:
        if name.endswith('/gamma'):
            scope = name[:-len('/gamma')]
            sc0 = w[name] / np.sqrt(w[f'{scope}/moving_variance'] + eps)
            sh0 = w[f'{scope}/beta'] - w[f'{scope}/moving_mean'] * sc0
            sc1 = m.weights[name] / np.sqrt(m.weights[f'{scope}/moving_variance'] + eps)
            sh1 = m.weights[f'{scope}/beta'] - m.weights[f'{scope}/moving_mean'] * sc1
            if scope.endswith('/BatchNorm'):                 # behind a convolution: the scale moved into the filter
                conv = scope[:-len('/BatchNorm')]
                assert np.allclose(sc1, 1.0, rtol=0, atol=2e-7)
                assert np.allclose(m.weights[f'{conv}/weights'], w[f'{conv}/weights'] * sc0, rtol=2e-6, atol=1e-9)
            else:
                assert np.allclose(sc1, sc0, rtol=2e-6, atol=0)
            assert np.allclose(sh1, sh0, rtol=2e-6, atol=1e-7)
    if stride == 32:
        img = synth_images(1, seed=3)
        a = OracleNet(spec, w, perm, 'fp64')(img)
        b = OracleNet(spec, dict(m.weights), perm, 'fp64')(img)
        assert np.abs(a - b).max() < 2e-3, np.abs(a - b).max()       # float32 rounding of the folded constants (measured 2e-4 mm)


def test_centered_stride_is_read_off_the_padding_attributes():
    """FLAGS.centered_stride is not stored in the graph; conv2d_same's SAME / explicit-pad + VALID choice per strided unit
    (resnet_utils.py:120-135) gives it away.  A --no-centered-stride export must not be loaded as a centred model."""
    ji = exported_joint_info('h36m')
    for centred in (True, False):
        for stride in (32, 16, 8):
            spec = NetSpec('resnet_v2_50', stride, 17, centered_stride=centred)
            w = synth_weights(spec, 1)
            m = import_frozen_graph(frozen_graph(spec, w, export_permutation('h36m'), list(ji.names), np.asarray(ji.edges)))
            assert m.centered_stride is centred and m.stride == stride
            assert [u.shift for u in m.spec.units] == [u.shift for u in spec.units]


def test_folded_convolution_renamed_by_the_transform_tool():
    """A folded convolution that took the name of the multiplication it absorbed is found through its filter constant,
    which keeps the variable's name."""
    spec = NetSpec('resnet_v2_50', 32, 17)
    w = synth_weights(spec, 2)
    ji = exported_joint_info('h36m')
    data = frozen_graph(spec, w, export_permutation('h36m'), list(ji.names), np.asarray(ji.edges), bn_form='folded',
                        fold_renames=True)
    m = import_frozen_graph(data)
    plain = import_frozen_graph(frozen_graph(spec, w, export_permutation('h36m'), list(ji.names), np.asarray(ji.edges),
                                             bn_form='folded'))
    assert list(m.weights) == list(plain.weights)
    for k in m.weights:
        assert np.array_equal(m.weights[k], plain.weights[k]), k
