"""repaq_b200 - B200-native FASTQ <-> .rfq chunk codec (byte-identical to OpenGene/repaq v0.5.1).

Python is only the binding: `Codec` mirrors the reference's RfqCodec surface (reference src/rfqcodec.h:17-43) and the
drivers `compress` / `decompress` / `compare` mirror Repaq::compress* / decompress* / compare* (src/repaq.cpp) on in-memory files; the
work is done by the CUDA library behind the C ABI of include/repaq_b200.h.
"""
from .codec import Codec, RepaqError, compare, compress, decompress, make_header  # noqa: F401
