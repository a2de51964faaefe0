from m_gat_graphsage_b200.nn import (GATConv, GCNConv, GINConv, Linear, MessagePassing, SAGEConv,  # noqa: F401
                                     global_add_pool, global_max_pool, global_mean_pool)
