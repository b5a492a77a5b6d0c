"""Joint tables baked into the exported model ('joint_names' / 'joint_edges' constants).

Restates ``JointInfo`` (src/data/datasets.py:52-109), the per-dataset tables
(src/data/h36m.py:25-31, src/data/mpi_inf_3dhp.py:21-32, src/data/datasets.py:142-154) and the
export-time permutation that moves the pelvis back to its dataset position
(src/main.py:119-128).  In this codebase the root (pelvis) is always the LAST model joint
(src/tfu3d.py:23-25).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple


def _pairwise(xs):
    return zip(xs[:-1], xs[1:])


def _edges_from_paths(names: Sequence[str], paths: str) -> List[Tuple[int, int]]:
    ids = {n: i for i, n in enumerate(names)}
    edges = []
    for path in paths.split(','):
        js = path.split('-')
        for a, b in _pairwise(js):
            if a in ids and b in ids:
                edges.append((ids[a], ids[b]))
    return edges


@dataclass
class JointInfo:
    names: List[str]
    edges: List[Tuple[int, int]]

    @property
    def n_joints(self) -> int:
        return len(self.names)

    @property
    def mirror_mapping(self) -> List[int]:
        """Index of the joint on the opposite body side (datasets.py:76-79,95-102): an 'l' / 'r' first letter is
        swapped, every other name maps to itself."""
        ids = {n: i for i, n in enumerate(self.names)}

        def other(name):
            if name.startswith('l'):
                return 'r' + name[1:]
            if name.startswith('r'):
                return 'l' + name[1:]
            return name
        return [ids[other(n)] for n in self.names]

    def permute_joints(self, permutation: Sequence[int]) -> 'JointInfo':
        """datasets.py:105-109.  As in the reference, edges whose endpoints are not selected by
        ``permutation`` cannot be expressed; the reference would raise on them, so do we."""
        inv = {old: new for new, old in enumerate(permutation)}
        names = [self.names[i] for i in permutation]
        edges = [(inv[i], inv[j]) for i, j in self.edges]
        return JointInfo(names, edges)


_H36M_NAMES = ('rhip,rkne,rank,lhip,lkne,lank,tors,neck,head,htop,'
               'lsho,lelb,lwri,rsho,relb,rwri,pelv').split(',')
_H36M_EDGES = ('htop-head-neck-lsho-lelb-lwri,neck-rsho-relb-rwri,'
               'neck-tors-pelv-lhip-lkne-lank,pelv-rhip-rkne-rank')

_TDHP_ALL = ('spi3,spi4,spi2,spin,pelv,neck,head,htop,lcla,lsho,lelb,lwri,lhan,rcla,rsho,relb,rwri,'
             'rhan,lhip,lkne,lank,lfoo,ltoe,rhip,rkne,rank,rfoo,rtoe').split(',')
_TDHP_SELECTED = [7, 5, 14, 15, 16, 9, 10, 11, 23, 24, 25, 18, 19, 20, 3, 6, 4]
_TDHP_EDGES = ('htop-head-neck-lsho-lelb-lwri,neck-rsho-relb-rwri,neck-spin-pelv-lhip-lkne-lank,'
               'pelv-rhip-rkne-rank')

_MERGED_NAMES = [
    'neck', 'nose', 'lsho', 'lelb', 'lwri', 'lhip', 'lkne', 'lank', 'rsho', 'relb',
    'rwri', 'rhip', 'rkne', 'rank', 'leye', 'lear', 'reye', 'rear', 'pelv',
    'htop_tdhp', 'neck_tdhp', 'rsho_tdhp', 'lsho_tdhp', 'rhip_tdhp', 'lhip_tdhp',
    'spin_tdhp', 'head_tdhp', 'pelv_tdhp', 'rhip_h36m', 'lhip_h36m', 'tors_h36m',
    'neck_h36m', 'head_h36m', 'htop_h36m', 'lsho_h36m', 'rsho_h36m', 'pelv_h36m',
    'lhip_tdpw', 'rhip_tdpw', 'bell_tdpw', 'che1_tdpw', 'che2_tdpw', 'ltoe_tdpw',
    'rtoe_tdpw', 'neck_tdpw', 'lcla_tdpw', 'rcla_tdpw', 'head_tdpw', 'lsho_tdpw',
    'rsho_tdpw', 'lhan_tdpw', 'rhan_tdpw', 'pelv_tdpw']
_MERGED_EDGES = [(1, 0), (0, 18), (0, 2), (2, 3), (3, 4), (0, 8), (8, 9), (9, 10), (18, 5), (5, 6),
                 (6, 7), (18, 11), (11, 12), (12, 13), (15, 14), (14, 1), (17, 16), (16, 1)]

# export permutations, src/main.py:119-125
_PERMUTATIONS = {
    'merged': [0, 1, 18, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17],
    'h36m': [16, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15],
    'mpi_inf_3dhp': [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 17, 14, 15],
}


def model_joint_info(dataset: str) -> JointInfo:
    """Joint set predicted by the head (model order, root last)."""
    if dataset == 'h36m':
        return JointInfo(list(_H36M_NAMES), _edges_from_paths(_H36M_NAMES, _H36M_EDGES))
    if dataset == 'mpi_inf_3dhp':
        names = [_TDHP_ALL[j] for j in _TDHP_SELECTED]
        return JointInfo(names, _edges_from_paths(names, _TDHP_EDGES))
    if dataset == 'merged':
        return JointInfo(list(_MERGED_NAMES), list(_MERGED_EDGES))
    if dataset == 'coco19':
        # BASELINE.json's "COCO/CMU 19 joints" head: the first 19 merged joints, pelvis last.
        return JointInfo(list(_MERGED_NAMES[:19]), list(_MERGED_EDGES))
    raise ValueError(f'unknown dataset {dataset!r}')


def export_permutation(dataset: str) -> List[int]:
    if dataset == 'coco19':
        return list(_PERMUTATIONS['merged'])
    if dataset == 'mpi_inf_3dhp':
        # src/main.py:125 lists index 17 for a 17-joint model (src/data/mpi_inf_3dhp.py:27):
        # tf.gather would fail on it, i.e. the reference cannot export this dataset as written.
        raise ValueError('mpi_inf_3dhp export permutation is out of range in the reference '
                         '(src/main.py:125 vs src/data/mpi_inf_3dhp.py:27)')
    return list(_PERMUTATIONS[dataset])


def exported_joint_info(dataset: str) -> JointInfo:
    """What the frozen graph's 'joint_names' / 'joint_edges' constants hold (main.py:128,140-141)."""
    return model_joint_info(dataset).permute_joints(export_permutation(dataset))
