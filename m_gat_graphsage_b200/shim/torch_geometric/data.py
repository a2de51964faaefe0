from m_gat_graphsage_b200.data import Batch, Data, DataLoader  # noqa: F401
