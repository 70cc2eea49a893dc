"""The multi-GPU call of the C-ABI (zpb_group_*): one archive, its entries cut into contiguous runs over every visible
GPU (two or more when the box has them; the same code path with one), results identical to the single-GPU call and to
the oracle.  The drop-in's batched read (zpack_read_files) goes through the same call."""
import numpy as np
import pytest

from zpack_b200 import container, corpus
import zpack_b200

pytestmark = pytest.mark.gpu


def _archive(oracle, n, size, method=2):
    bufs = [corpus.entry_bytes(i, size if i % 7 else size // 3) for i in range(n)]
    frames = [oracle.lz4f_encode_port(b, 0, independent=bool(i & 1)) for i, b in enumerate(bufs)]
    hashes = [oracle.xxh3_port(b) for b in bufs]
    arch = container.assemble([corpus.entry_name(i) for i in range(n)], frames, [len(b) for b in bufs], hashes, [method] * n)
    return bufs, hashes, arch


def test_group_unpack_over_all_visible_gpus(oracle):
    import torch
    ngpu = torch.cuda.device_count()
    bufs, hashes, arch = _archive(oracle, 96, 131072)
    e = container.parse(arch).entries()
    out_size = int((e["dst_off"] + e["dst_cap"]).max())
    for devices in ([0], list(range(ngpu)), None):
        g = zpack_b200.Group(devices)
        assert g.size == (ngpu if devices is None else len(devices))
        out = np.zeros(out_size + 16, np.uint8)
        status, digest = g.unpack_host(arch, len(arch), out, out_size, e)
        assert (status == 0).all(), status
        assert np.array_equal(digest, np.array(hashes, np.uint64))
        for i, b in enumerate(bufs):
            o = int(e["dst_off"][i])
            assert np.array_equal(out[o:o + len(b)], b), i
        g.close()


def test_group_reports_per_entry_errors_like_the_single_gpu_call(oracle, gpu_ctx):
    bufs, hashes, arch = _archive(oracle, 40, 70000)
    d = container.parse(arch)
    e = d.entries()
    e["hash"][3] ^= 1                       # FILE_HASH_MISMATCH
    e["dst_cap"][5] = 10                    # BUFFER_TOO_SMALL
    e["src_off"][7] = len(arch) + 5         # FILE_OFFSET_INVALID
    e["method"][9] = 77                     # COMP_METHOD_INVALID
    out_size = int((e["dst_off"] + np.maximum(e["dst_cap"], e["uncomp_size"])).max())
    out1 = np.zeros(out_size + 16, np.uint8)
    st1, dg1 = gpu_ctx.unpack_host(arch, len(arch), out1, out_size, e)
    g = zpack_b200.Group(None)
    out2 = np.zeros(out_size + 16, np.uint8)
    st2, dg2 = g.unpack_host(arch, len(arch), out2, out_size, e)
    g.close()
    assert list(st1) == list(st2)
    assert st2[3] == 15 and st2[5] == 12 and st2[7] == 16 and st2[9] == 19
    ok = st2 == 0
    assert np.array_equal(dg1[ok], dg2[ok])
    for i in np.flatnonzero(ok):            # slots of failed entries hold whatever the device buffer held: unspecified, as in the reference
        o, n = int(e["dst_off"][i]), len(bufs[i])
        assert np.array_equal(out1[o:o + n], bufs[i]) and np.array_equal(out2[o:o + n], bufs[i]), i


@pytest.mark.parametrize("method", [2, 1])
def test_group_pack_round_trips(oracle, method):
    from zpack_b200 import lib as zlib
    g = zpack_b200.Group(None)
    ctx = zpack_b200.Context(0)
    sizes = [131072] * 40 + [0, 1, 70000]
    bufs = [corpus.entry_bytes(i, s) for i, s in enumerate(sizes)]
    f = np.zeros(len(bufs), zlib.File)
    in_off = out_off = 0
    for i, b in enumerate(bufs):
        cap = ctx.pack_bound(method, len(b))
        f["src_off"][i], f["size"][i], f["dst_off"][i], f["dst_cap"][i], f["method"][i] = in_off, len(b), out_off, cap, method
        in_off += (len(b) + 15) & ~15
        out_off += (cap + 15) & ~15
    h_in = np.zeros(in_off + 16, np.uint8)
    for i, b in enumerate(bufs):
        h_in[int(f["src_off"][i]):int(f["src_off"][i]) + len(b)] = b
    h_out = np.zeros(out_off + 16, np.uint8)
    comp, digest, status = g.pack_host(h_in, len(h_in), h_out, len(h_out), f)
    assert (status == 0).all()
    for i, b in enumerate(bufs):
        fr = h_out[int(f["dst_off"][i]):int(f["dst_off"][i] + comp[i])]
        rc, got = (oracle.lz4f_decode_port if method == 2 else oracle.zstd_decode_port)(fr, len(b))
        assert rc == 0 and np.array_equal(got[:len(b)], b) and int(digest[i]) == oracle.xxh3_port(b)
    g.close(); ctx.close()
