from m_gat_graphsage_b200.explain import ExplainerConfig, ModelConfig  # noqa: F401
