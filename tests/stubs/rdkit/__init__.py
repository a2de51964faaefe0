"""TEST INFRASTRUCTURE: a synthetic stand-in for RDKit (which cannot be installed here), just large enough for the
featurisers of the reference scripts (train.py:25-63, test.py:20-58, ablation/model1.py) to run UNCHANGED.

A "SMILES" string of the form ``SYN<seed>`` denotes the synthetic molecule ``synth_batch(1, seed)``; ``MolFromSmiles``
returns an object whose atoms answer ``GetSymbol / GetDegree / GetImplicitValence / GetHybridization / GetIsAromatic /
GetTotalNumHs`` with the values that molecule's one-hot features encode, so the script's own featuriser reproduces the
generator's ``x`` and ``edge_index`` exactly.  Anything else parses to ``None`` like an invalid SMILES."""
__version__ = "0.0+synthetic"
from . import Chem  # noqa: F401
