from virtex_b200.modules import WordAndPositionalEmbedding  # noqa: F401
