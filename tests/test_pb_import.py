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
