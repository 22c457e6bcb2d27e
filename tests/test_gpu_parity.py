"""Parity tests proper: the CUDA library on a real GPU, through the C ABI, against the reference's golden vectors
and the oracle, plus size-independent properties at larger sizes.  Run with `-m gpu` on the B200 box."""
import ctypes as C
import os

import numpy as np
import pytest

from repaq_b200 import _lib
from repaq_b200 import codec as K
from tests import parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def codec():
    cd = K.Codec(device=0)          # the in-tree CUDA build; raises if it is missing or there is no GPU
    yield cd
    cd.close()


@pytest.mark.parametrize("name", parity.OK_CASES)
def test_encode_golden(codec, name):
    parity.check_encode_golden(codec, name)


@pytest.mark.parametrize("name", parity.ERR_CASES)
def test_encode_error(codec, name):
    parity.check_encode_error(codec, name)


@pytest.mark.parametrize("name", parity.OK_CASES)
def test_decode_golden(codec, name):
    parity.check_decode_golden(codec, name)


@pytest.mark.parametrize("name", sorted(parity.RLE_MAN))
def test_decode_run_length_quality_golden(codec, name):
    parity.check_decode_rle_golden(codec, name)


def test_config0_se_50k_reads(codec):
    """BASELINE.json configs[0]: single-end 50k-read 150bp, encode == reference algorithm, decode restores the input"""
    from tools import fqgen
    r1, _ = fqgen.generate(50000, seed=1)
    parity.check_against_oracle(codec, r1)


@pytest.mark.parametrize("flags,label", [(0, "plain"), (1, "npos"), (4, "varlen"), (8, "300bp"), (2, "crlf"), (12, "300bp-varlen")])
def test_pe_shapes_against_oracle(codec, flags, label):
    from tools import fqgen
    r1, r2 = fqgen.generate(60000, seed=40 + flags, paired=True, flags=flags)
    parity.check_against_oracle(codec, r1, r2, roundtrip=not (flags & 2))


def test_bgi_shape_against_oracle(codec):
    from tools import fqgen
    r1, _ = fqgen.generate(120000, seed=5, shape=fqgen.BGI)
    parity.check_against_oracle(codec, r1)


def test_small_chunks_and_tile_changes(codec):
    from tools import fqgen
    r1, r2 = fqgen.generate(30000, seed=51, paired=True, first_row=1950)       # crosses a tile boundary
    parity.check_against_oracle(codec, r1, r2, k=100)
    parity.check_against_oracle(codec, r1, r2, k=333)


def test_batches_concatenate(codec):
    """Streaming property: encoding a file in batches cut at the reported resume offsets gives the same bytes."""
    from tools import fqgen
    r1, r2 = fqgen.generate(45000, seed=52, paired=True)
    h = K.make_header(r1, r2)
    codec.set_header(h)
    whole, infos, _ = codec.encode(r1, r2)
    # first batch: only the first ~40 % of each file, not final -> whole chunks only
    a1, a2 = r1[: int(r1.size * 0.4)], r2[: int(r2.size * 0.4)]
    p1, i1, m1 = codec.encode(a1, a2, final=False)
    assert 0 < len(i1) < len(infos)
    p2, i2, m2 = codec.encode(r1[m1["r1_consumed"]:], r2[m1["r2_consumed"]:], final=True)
    assert p1 + p2 == whole


def test_device_pointer_api(codec):
    """mem=RPQ_MEM_DEVICE in and out: torch only provides the device memory"""
    import torch
    from tools import fqgen
    r1, r2 = fqgen.generate(30000, seed=53, paired=True)
    h = K.make_header(r1, r2)
    codec.set_header(h)
    host, _, _ = codec.encode(r1, r2)
    t1 = torch.from_numpy(r1.copy()).cuda()
    t2 = torch.from_numpy(r2.copy()).cuda()
    out = codec.encode_raw(t1.data_ptr(), t1.numel(), t2.data_ptr(), t2.numel(), 1, False, 1000000, True, (K.NEVER, K.NEVER), 0, 1)
    dev = torch.empty(out.bytes, dtype=torch.uint8, device="cuda")
    C.cdll.LoadLibrary("libcudart.so.12").cudaMemcpy(C.c_void_p(dev.data_ptr()), C.c_void_p(out.data), C.c_size_t(out.bytes), 3)
    assert bytes(dev.cpu().numpy()) == host
    # decode from device memory into device memory
    dec = K.Codec(device=0)
    dec.set_header(h)
    o = dec.decode_raw(dev.data_ptr(), dev.numel(), 1, True, 1)
    assert (o.out1_bytes, o.out2_bytes) == (r1.size, r2.size)
    g1 = torch.empty(o.out1_bytes, dtype=torch.uint8, device="cuda")
    C.cdll.LoadLibrary("libcudart.so.12").cudaMemcpy(C.c_void_p(g1.data_ptr()), C.c_void_p(o.out1), C.c_size_t(o.out1_bytes), 3)
    assert np.array_equal(g1.cpu().numpy(), r1)
    dec.close()


def test_large_roundtrip_property(codec):
    """~1 GB paired-end NovaSeq-shape: decode(encode(x)) == x, chunk sizes add up, and the first 60 MB of the encoding
    equal the oracle's (a checksum of the whole file against the oracle is bench.py's job at full size)."""
    from oracle import oracle as O
    from tools import fqgen
    r1, r2 = fqgen.generate(1400000, seed=54, paired=True)
    rfq = K.compress(r1, r2, codec=codec)
    d1, d2 = K.decompress(rfq, pe_out=True, codec=codec)
    assert np.array_equal(np.frombuffer(d1, dtype=np.uint8), r1) and np.array_equal(np.frombuffer(d2, dtype=np.uint8), r2)
    # prefix check against the oracle: the first 25 chunks only depend on the first 25 * 3334 pairs
    n_pairs = 25 * 3334
    from tools.fqgen import truncate_reads
    p1, p2 = truncate_reads(r1, n_pairs), truncate_reads(r2, n_pairs)
    ref = O.compress(bytes(p1), bytes(p2))
    assert rfq[: len(ref)] == ref


class _TorchDevice:
    """torch only provides the device memory"""
    @staticmethod
    def put(arr):
        import torch
        t = torch.from_numpy(arr.copy()).cuda()
        return t, t.data_ptr()

    @staticmethod
    def get(ptr, n):
        if not n:
            return b""
        host = np.empty(n, dtype=np.uint8)
        rc = C.cdll.LoadLibrary("libcudart.so.12").cudaMemcpy(C.c_void_p(host.ctypes.data), C.c_void_p(ptr), C.c_size_t(n), 2)
        assert rc == 0
        return host.tobytes()


def test_parallel_chunk_walk(monkeypatch):
    from tests import test_emu_parity as E
    E.test_parallel_chunk_walk(monkeypatch, lib_path=None, mem=_TorchDevice)


def test_dense_quality_spans(monkeypatch):
    """k_streams7 on the GPU: dense inputs against the oracle (default hand-over from k_streams4) and every quality span of ordinary
    golden inputs (RPQ_DEBUG_STREAMS5=2)"""
    from tests import test_emu_parity as E
    for knob in ("1", "2"):
        monkeypatch.setenv("RPQ_DEBUG_STREAMS5", knob)
        cd = K.Codec(device=0)
        try:
            for name in ("nova_pe_k1000", "bgi_se_varlen_k100", "nova_se_late_quality", "nova_pe_k100_npos"):
                parity.check_encode_golden(cd, name)
            E._dense_cases(cd, None)
            if knob == "1":
                E.test_dense_hint_follows_the_data(cd)
        finally:
            cd.close()


def test_window_cut_and_blank_lines(codec):
    from tests import test_emu_parity as E
    E.test_window_cut_inside_the_chunk_closing_record(codec)
    E.test_blank_lines(codec)
    E.test_crlf_on_reader_buffer_edges(codec)
    E.test_lone_cr_and_mixed_line_ends(codec)


def test_library_really_ran_on_gpu(codec):
    parity.check_encode_golden(codec, "kat_pe")              # the statistics are those of the last call
    s = codec.stats()
    assert s.launches > 0
    assert os.path.basename(_lib.LIB_PATH) == "librepaq_b200.so"


def test_n_positions_in_few_reads(codec):
    parity.check_n_positions_in_few_reads(codec, n_pairs=60000)


def test_adversarial_quality_columns(codec):
    parity.check_adversarial_quality_columns(codec, n_reads=60000)


def test_control_bytes_in_names(codec):
    parity.check_control_bytes_in_names(codec)


def test_read_longer_than_the_header_can_store(codec):
    """the header's read length width comes from the first chunk (src/rfqcodec.cpp:48-53): a later read of more than 255 bases is refused"""
    short = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, b"ACGT" * 25, b"F" * 100) for i in range(1200))
    long_ = b"@long\n%s\n+\n%s\n" % (b"ACGT" * 150, b"F" * 600)
    with pytest.raises(K.RepaqError) as e:
        K.compress(short + long_, k=100, codec=codec)
    assert "does not fit the header" in str(e.value)
    parity.check_against_oracle(codec, long_ + short, k=100)


@pytest.mark.parametrize("reads", ["64", "32"])
def test_small_formatter_tiles(monkeypatch, reads):
    monkeypatch.setenv("RPQ_DEBUG_FMT_READS", reads)
    cd = K.Codec(device=0)
    try:
        for name in ("nova_pe_k1000", "nova_pe_k100_npos", "bgi_se_varlen_k100", "nova_pe_300bp_varlen_k100"):
            parity.check_decode_golden(cd, name)
    finally:
        cd.close()


def test_quality_longer_than_sequence(codec):
    parity.check_quality_longer_than_sequence(codec)


def test_crlf_on_reader_buffer_edges(codec):
    """the reference's reader at its 1 MiB refills, with and without a final line break (Q13 on the flush chunk): the CPU test's cases"""
    from tests.test_emu_parity import test_crlf_on_reader_buffer_edges as cases
    cases(codec)


def test_medium_density_quality_columns(codec):
    parity.check_medium_density(codec)
