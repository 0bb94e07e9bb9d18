from virtex_b200.modules import TextualHead, TransformerDecoderTextualHead  # noqa: F401
