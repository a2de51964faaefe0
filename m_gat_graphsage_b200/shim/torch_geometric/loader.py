from m_gat_graphsage_b200.data import DataLoader  # noqa: F401
