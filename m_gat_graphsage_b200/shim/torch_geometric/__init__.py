"""``torch_geometric``-shaped import shim over ``m_gat_graphsage_b200``.

Put this directory's parent on ``sys.path`` (``python -m m_gat_graphsage_b200.run script.py`` does) and
the reference scripts' ``from torch_geometric.nn import GATConv, SAGEConv, global_max_pool`` /
``from torch_geometric.data import Data, DataLoader`` resolve to the sm_100a implementations.
Only the names the reference imports exist here."""
__version__ = "2.6.1+mgs_b200"

from . import data, explain, loader, nn  # noqa: F401
