"""Layer-plan replay (metro_pose3d_b200/spec.py and csrc/plan.cpp) against the hand-derived tables of
SURVEY.md section 8a (which follow resnet_utils.py:307-350 and resnet_v2.py:277-302)."""
import pytest

from metro_pose3d_b200 import lib
from metro_pose3d_b200.spec import NetSpec, same_pad


EXPECT = {  # (arch, stride): (n_convs, GFLOP/crop, sides after b1..b4, conv2 rates of first unit of b1..b4)
    ('resnet_v2_50', 32): (54, 9.127, (32, 16, 8, 8), (1, 1, 1, 1)),
    ('resnet_v2_50', 16): (54, 15.299, (32, 16, 16, 16), (1, 1, 1, 2)),
    ('resnet_v2_50', 8): (54, 49.944, (32, 32, 32, 32), (1, 1, 2, 4)),
    ('resnet_v2_101', 16): (105, 25.013, (32, 16, 16, 16), (1, 1, 1, 2)),
    ('resnet_v2_101', 4): (105, 350.080, (64, 64, 64, 64), (1, 2, 4, 8)),
}


def _j(stride, arch):
    return 17 if (arch, stride) in (('resnet_v2_50', 32), ('resnet_v2_50', 16)) else 19


@pytest.mark.parametrize('key', list(EXPECT))
def test_plan_tables(key):
    arch, stride = key
    sp = NetSpec(arch, stride, _j(stride, arch))
    n_convs, gflop, sides, rates = EXPECT[key]
    assert len(sp.convs) == n_convs
    assert abs(sp.flops_per_crop / 1e9 - gflop) < 5e-4
    got_sides, got_rates = [], []
    for b in range(1, 5):
        units = [u for u in sp.units if u.name.startswith(f'block{b}/')]
        got_sides.append(units[-1].out_side)
        got_rates.append(units[0].rate)
    assert tuple(got_sides) == sides
    assert tuple(got_rates) == rates
    assert sp.root.out_side == 128 and sp.pool_out == 64 and sp.feat_side == 256 // stride


def test_centered_stride_block_selection():
    # Q5: rn50 S=32 -> block3, 16 -> block2, 8 -> block1, 4 -> none; rn101 S=4 sets block3 but inert
    for arch in ('resnet_v2_50', 'resnet_v2_101'):
        for stride, blk in ((32, 'block3'), (16, 'block2'), (8, 'block1'), (4, None)):
            sp = NetSpec(arch, stride, 17)
            shifted = [u.name.split('/')[0] for u in sp.units if u.shift]
            assert shifted == ([blk] if blk else []), (arch, stride, shifted)
            for u in sp.units:
                if u.shift:
                    assert (u.conv2.pad_lo, u.conv2.pad_hi, u.stride) == (0, 1, 2)   # TF SAME, even input
                elif u.stride == 2:
                    assert (u.conv2.pad_lo, u.conv2.pad_hi) == (1, 1)                # explicit pad (Q4)
                else:
                    assert u.conv2.pad_lo == u.conv2.pad_hi == u.rate                # SAME, stride 1
    sp = NetSpec('resnet_v2_50', 16, 17, centered_stride=False)
    assert not any(u.shift for u in sp.units)


def test_projection_only_in_first_units_and_never_strided():
    for arch in ('resnet_v2_50', 'resnet_v2_101'):
        for stride in (4, 8, 16, 32):
            sp = NetSpec(arch, stride, 19)
            for u in sp.units:
                assert (u.shortcut is not None) == u.name.endswith('unit_1')
                if u.shortcut is not None:
                    assert u.stride == 1


@pytest.mark.parametrize('stride', [3, 6, 10, 64])
def test_bad_stride_raises_value_error(stride):
    # resnet_v2.py:213-214, resnet_utils.py:333,345,348
    with pytest.raises(ValueError):
        NetSpec('resnet_v2_50', stride, 17)
    with pytest.raises(ValueError):
        lib.plan_describe(lib.make_spec('resnet_v2_50', stride, 17, [0]))


def test_same_pad_formula():
    assert same_pad(64, 3, 2) == (32, 0, 1)
    assert same_pad(64, 3, 1) == (64, 1, 1)
    assert same_pad(64, 17, 1) == (64, 8, 8)
    assert same_pad(65, 3, 2) == (33, 1, 1)


@pytest.mark.parametrize('arch,stride,j', [('resnet_v2_50', 32, 17), ('resnet_v2_50', 16, 17), ('resnet_v2_50', 8, 19),
                                           ('resnet_v2_101', 16, 19), ('resnet_v2_101', 4, 19), ('resnet_v2_50', 4, 53)])
def test_cpp_plan_matches_python(libmetro, arch, stride, j):
    """The C-ABI library replays the graph on its own (csrc/plan.cpp); both replays must agree."""
    from metro_pose3d_b200.weights import blob_order, blob_size
    sp = NetSpec(arch, stride, j)
    d = lib.plan_describe(lib.make_spec(arch, stride, j, list(range(j))))
    assert d['n_convs'] == len(sp.convs)
    assert d['blob_floats'] == blob_size(sp)
    assert abs(d['flops_per_crop'] - sp.flops_per_crop) < 1
    offs, off = {}, 0
    for name, shape in blob_order(sp):
        offs[name] = off
        n = 1
        for s in shape:
            n *= s
        off += n
    for c_py, c_cc in zip(sp.convs, d['convs']):
        for f in ('name', 'cin', 'cout', 'k', 'stride', 'rate', 'pad_lo', 'pad_hi', 'in_side', 'out_side',
                  'has_bias', 'has_bn', 'relu'):
            assert getattr(c_py, f) == c_cc[f], (c_py.name, f)
        scope = c_py.name if c_py.name in ('conv1', 'logits') else c_py.name.replace('/', '/bottleneck_v2/', 2).replace(
            '/bottleneck_v2/', '/', 1)
        assert offs[scope + '/weights'] == c_cc['w_off'], c_py.name
    for u_py, u_cc in zip(sp.units, d['units']):
        for f in ('name', 'cin', 'depth', 'cb', 'stride', 'rate', 'shift', 'in_side', 'out_side'):
            assert getattr(u_py, f) == u_cc[f]
        assert (u_py.shortcut is not None) == u_cc['proj']
