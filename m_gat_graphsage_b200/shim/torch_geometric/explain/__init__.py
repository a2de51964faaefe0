from m_gat_graphsage_b200.explain import Explainer, Explanation, GNNExplainer  # noqa: F401

from . import config  # noqa: F401
