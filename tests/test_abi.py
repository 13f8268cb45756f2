"""CPU: the C-ABI library loads, exports every symbol include/neo_b200.h declares, and -- with no GPU -- refuses to
compute instead of falling back to anything."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "neo_b200.h")).read()
    return sorted(set(re.findall(r"NEO_B200_API[^;]*?\b(neo_b200_\w+)\s*\(", text)))


def test_header_and_binding_agree(pkg):
    names = declared_symbols()
    assert len(names) >= 38
    assert sorted(pkg.SIGNATURES) == names


def test_config_struct_layout_matches_the_header(pkg):
    # neo_b200_conv_config is passed by pointer: the ctypes mirror must list the same fields, in order, with the same widths
    text = open(os.path.join(ROOT, "include", "neo_b200.h")).read()
    body = re.search(r"typedef struct neo_b200_conv_config\s*\{(.*?)\}\s*neo_b200_conv_config;", text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\b(int|size_t)\s+(\w+)\s*;", body)
    want = [(name, ctypes.c_int if ctype == "int" else ctypes.c_size_t) for ctype, name in fields]
    assert [(n, t) for n, t in pkg.ConvConfig._fields_] == want
    assert "frame_blocks" in [n for n, _ in want]


def test_library_exports_every_declared_symbol(pkg):
    lib = ctypes.CDLL(pkg.LIBRARY_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_integer_helpers_need_no_device(pkg):
    assert pkg.FFTPlan.max_order() == 27  # c2c_dit2_plan.hpp:59-62
    assert pkg.num_partitions(4096, 128) == 32 and pkg.num_partitions(4095, 128) == 32  # uniform_partition_test.cpp
    assert pkg.num_partitions(1 << 20, 1024) == 1024
    assert [pkg.next_order(n) for n in (1, 2, 3, 1023, 1024, 1025)] == [0, 1, 2, 10, 10, 11]
    assert pkg.library().neo_b200_version().startswith(b"neo_b200")


def test_rfftfreq_helper(pkg):
    import numpy as np

    # fft/rfftfreq.hpp:12-29: index * fs / size over `size` entries
    got = pkg.rfftfreq(8, 1.0 / 48000.0)
    assert got.dtype == np.float32 and np.allclose(got, np.arange(8) * 48000.0 / 8)
    assert np.allclose(pkg.rfftfreq(5, 0.5, "float64"), np.arange(5) * 2.0 / 5)


def test_order_past_max_is_unsupported_like_the_reference(pkg):
    # fft_test.cpp:62-67 expects a throw for next_order(max_size()+1)
    with pytest.raises(RuntimeError, match="unsupported order"):
        pkg.FFTPlan(pkg.FFTPlan.max_order() + 1)
    with pytest.raises(RuntimeError, match="unsupported order"):
        pkg.RFFTPlan(28)


def test_convolver_configuration_is_validated_before_any_device_work(pkg):
    # neo_b200_conv_create checks its configuration first, so the error contract is testable without a GPU
    def create(**kw):
        cfg = dict(kind=0, dtype=0, topology=0, outputs=2, inputs=2, block=64, partitions=8, max_blocks=0, partition_begin=0,
                   partition_end=0, frame_blocks=0, input_delayed=0)
        cfg.update(kw)
        h = ctypes.c_void_p()
        c = pkg.ConvConfig(*[cfg[name] for name, _ in pkg.ConvConfig._fields_])
        status = pkg.library().neo_b200_conv_create(ctypes.byref(h), ctypes.byref(c))
        message = pkg.library().neo_b200_last_error().decode()
        if status == 0:
            pkg.library().neo_b200_conv_destroy(h)
        return status, message

    for kw, needle in (
        (dict(block=96), "power of two"),
        (dict(partitions=0), "partitions"),
        (dict(frame_blocks=3), "frame_blocks"),
        (dict(frame_blocks=1024), "frame_blocks"),
        (dict(frame_blocks=4, max_blocks=8), "frame_blocks"),
        (dict(frame_blocks=4, partition_begin=2, partition_end=8), "multiple of frame_blocks"),
        (dict(partition_begin=5, partition_end=3), "partition range"),
        (dict(topology=0, inputs=3), "diagonal"),
        (dict(kind=7), "kind"),
    ):
        status, message = create(**kw)
        assert status != 0 and needle in message, (kw, message)
    status, message = create(frame_blocks=4)  # a valid configuration gets as far as the device check
    assert status == 0 or "no CPU fallback" in message


def test_bank_layout_arithmetic_needs_no_device(pkg):
    # neo_b200_bank_layout_info: rank = group * shards + shard; rows and partitions per rank; frame-aligned shards with their delay
    info = lambda layout, rank, frame=256: pkg.Bank.layout_info(pkg.UPOLS, "float32", pkg.DIAGONAL, 1024, 1024, 1024, 1024, frame or 16, frame, layout, rank)
    r = info((4, 2), 5)
    assert (r["channel_group"], r["partition_shard"]) == (2, 1)
    assert (r["group_first"], r["group_count"]) == (512, 256)
    assert (r["in_first"], r["in_count"], r["out_first"], r["out_count"]) == (640, 128, 640, 128)
    assert (r["partition_begin"], r["partition_end"], r["delay_blocks"]) == (512, 1024, 512)
    r = info((1, 8), 3, frame=0)  # direct form: no delayed input, the handle reaches back through its own ring
    assert (r["partition_begin"], r["partition_end"], r["delay_blocks"]) == (384, 512, 0)
    covered = []
    for rank in range(8):
        r = info((2, 4), rank)
        covered.append((r["channel_group"], r["partition_begin"], r["partition_end"]))
        assert r["partition_begin"] % 256 == 0
    assert covered == [(g, s * 256, (s + 1) * 256) for g in range(2) for s in range(4)]
    for bad in ((3, 2), (1, 16)):
        with pytest.raises(RuntimeError):
            pkg.Bank.layout_info(pkg.UPOLS, "float32", pkg.DIAGONAL, 1024, 1024, 1024, 1024, 256, 256, bad, 0)
    with pytest.raises(RuntimeError, match="shards"):  # 1024 partitions = 4 frames of 256: cannot be cut into 8 shards
        info((1, 8), 0)


def test_no_cpu_fallback(pkg):
    if pkg.device_count() > 0:
        pytest.skip("a CUDA device is present")
    for make in (lambda: pkg.FFTPlan(4), lambda: pkg.RFFTPlan(4), lambda: pkg.bitrev_table(4)):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            make()


def test_product_never_touches_the_oracle():
    for base, _, files in os.walk(os.path.join(ROOT, "neo-dsp_b200")):
        if os.sep + "build" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")) or f == "Makefile":
                text = open(os.path.join(base, f)).read()
                assert "oracle" not in text.lower() or f == "__init__.py" and "pyoracle" not in text, f
    text = open(os.path.join(ROOT, "include", "neo_b200.h")).read()
    assert "oracle" not in text.lower()
