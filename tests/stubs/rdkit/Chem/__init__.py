import enum
import re

import numpy as np


class _HybridizationType(enum.Enum):
    UNSPECIFIED = 0
    S = 1
    SP = 2
    SP2 = 3
    SP3 = 4
    SP3D = 5
    SP3D2 = 6
    OTHER = 7


class rdchem:  # noqa: N801  (module-like namespace: Chem.rdchem.HybridizationType.SP ...)
    HybridizationType = _HybridizationType


_SYMBOLS = ["C", "N", "O", "S", "F", "P", "Cl", "Br", "I", "Si"]          # the 10th is outside the featuriser's list
_HYBRID = [_HybridizationType.SP, _HybridizationType.SP2, _HybridizationType.SP3, _HybridizationType.SP3D,
           _HybridizationType.SP3D2]


class Atom:
    def __init__(self, feat):
        f = [int(v) for v in feat]
        self._symbol = _SYMBOLS[f[0:10].index(1)]
        self._degree = f[10:17].index(1)
        self._valence = f[17:24].index(1)
        self._hybrid = _HYBRID[f[24:29].index(1)] if 1 in f[24:29] else _HybridizationType.OTHER
        self._aromatic = bool(f[29])
        self._hs = f[30:35].index(1)

    def GetSymbol(self): return self._symbol  # noqa: E704, N802
    def GetDegree(self): return self._degree  # noqa: E704, N802
    def GetImplicitValence(self): return self._valence  # noqa: E704, N802
    def GetHybridization(self): return self._hybrid  # noqa: E704, N802
    def GetIsAromatic(self): return self._aromatic  # noqa: E704, N802
    def GetTotalNumHs(self): return self._hs  # noqa: E704, N802


class Bond:
    def __init__(self, a, b):
        self._a, self._b = int(a), int(b)

    def GetBeginAtomIdx(self): return self._a  # noqa: E704, N802
    def GetEndAtomIdx(self): return self._b  # noqa: E704, N802


class Mol:
    def __init__(self, seed):
        from m_gat_graphsage_b200.synth import synth_batch
        b = synth_batch(1, seed, device="cpu")
        self.seed = seed
        self._atoms = [Atom(row) for row in b.x.tolist()]
        ei = b.edge_index.tolist()
        self._bonds = [Bond(i, j) for i, j in zip(ei[0], ei[1]) if i < j]

    def GetNumAtoms(self): return len(self._atoms)  # noqa: E704, N802
    def GetAtoms(self): return list(self._atoms)  # noqa: E704, N802
    def GetBonds(self): return list(self._bonds)  # noqa: E704, N802


def MolFromSmiles(smiles):  # noqa: N802
    m = re.fullmatch(r"SYN(\d+)", str(smiles).strip())
    return Mol(int(m.group(1))) if m else None


class _AllChem:
    @staticmethod
    def GetMorganFingerprintAsBitVect(mol, radius, nBits=2048):  # noqa: N802, N803
        rng = np.random.default_rng(1_000_003 * mol.seed + radius)
        return (rng.random(nBits) < 0.05).astype(np.int64).tolist()


AllChem = _AllChem()
