"""Host-side logic of the slab decomposition with world_size = 2 over gloo (no GPU): rendezvous + NCCL unique-id
broadcast plumbing, slab partitioning, and the exchange layout of csrc/fft_plan.cu::exec_dist restated in NumPy
(y-pass output written destination-rank-major [peer][kx, y_local, z_local]; blocks received into (nkr, ny/P, nz)), and the
blocked receive layouts of the fused pass + collective exchange (peer stores emulated by shipping (address, value) lists)."""
import os
import socket

import numpy as np
import pytest

import oracle  # noqa: F401  (tests may use the oracle)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import scipy.fft as sfft
    import torch.distributed as dist
    import fourierflows_jl_b200 as ff
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. the plumbing bench.py / Dist.from_torch use: rank 0 creates the 128-byte id, everyone receives the same bytes
        obj = [ff.Dist.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        ids = [None] * world
        dist.all_gather_object(ids, obj[0])
        assert all(i == ids[0] for i in ids) and len(ids[0]) == 128
        # 2. forward slab transform restated with the C code's layout, exchange over gloo
        nx, ny, nz = 16, 8, 12
        P = world
        rng = np.random.default_rng(0)
        x = np.asfortranarray(rng.standard_normal((nx, ny, nz)))
        ref = sfft.rfftn(x, axes=(2, 1, 0))
        nkr, nyl, nzl = nx // 2 + 1, ny // P, nz // P
        xl = ff.physical_slab(x, P, rank)
        a = sfft.fft(sfft.rfft(xl, axis=0), axis=1)                       # local x r2c + y c2c: (nkr, ny, nzl)
        send = np.zeros((P, nkr, nyl, nzl), dtype=complex)                # [peer][kx, y_local, z_local]
        for y in range(ny):
            send[y // nyl, :, y % nyl, :] = a[:, y, :]                    # the segmented output stride of the y-pass
        blocks = [None] * P
        for src in range(P):
            got = [None] * P if rank == src else None
            # all-to-all via gather/scatter objects (gloo has no all_to_all)
            gathered = [None] * P
            dist.all_gather_object(gathered, send[src])
            if rank == src:
                blocks = gathered
        recv = np.zeros((nkr, nyl, nz), dtype=complex, order="F")
        for s in range(P):
            recv[:, :, s * nzl:(s + 1) * nzl] = blocks[s]                 # block from rank s lands at z = s*nzl + z_local
        out = sfft.fft(recv, axis=2)
        err = np.linalg.norm(out - ff.spectral_slab(ref, P, rank)) / np.linalg.norm(ref)
        assert err < 1e-13, err
        # 2b. the fused pass + collective (FFB_EXCHANGE_PEER_STORE): the y-pass stores into the destination rank's receive buffer
        #     with the blocked layout [kx block][z][y_local][B] of exec_dist (B complex = 64 bytes; ragged last kx block), the
        #     z-pass reads that layout; inverse: [kx block][y][z_local][B].  Peer stores are emulated by shipping (address, value).
        B = 4
        nkt = (nkr + B - 1) // B
        kx, yy, zl = np.meshgrid(np.arange(nkr), np.arange(ny), np.arange(nzl), indexing="ij")
        addr = (((kx // B) * nz + (rank * nzl + zl)) * nyl + (yy % nyl)) * B + kx % B
        stores = [(addr[yy // nyl == q], a[yy // nyl == q]) for q in range(P)]
        inbox = [None] * P
        for dst in range(P):
            gathered = [None] * P
            dist.all_gather_object(gathered, stores[dst])
            if rank == dst:
                inbox = gathered
        flat = np.full(nkt * nz * nyl * B, np.nan + 0j)
        for ad, val in inbox:
            flat[ad] = val
        kx2, yl2, z2 = np.meshgrid(np.arange(nkr), np.arange(nyl), np.arange(nz), indexing="ij")
        recv_b = flat[(((kx2 // B) * nz + z2) * nyl + yl2) * B + kx2 % B]          # what the z-pass loads
        assert np.array_equal(recv_b, recv)                                        # same data as the NCCL-layout exchange
        spec_l = sfft.fft(recv_b, axis=2)
        assert np.linalg.norm(spec_l - ff.spectral_slab(ref, P, rank)) / np.linalg.norm(ref) < 1e-13
        # inverse: z-pass stores into rank (z // nzl) at [kx block][y = rank*nyl + y_local][z_local][B]
        bz = sfft.ifft(spec_l, axis=2)
        addr = (((kx2 // B) * ny + (rank * nyl + yl2)) * nzl + (z2 % nzl)) * B + kx2 % B
        stores = [(addr[z2 // nzl == q], bz[z2 // nzl == q]) for q in range(P)]
        for dst in range(P):
            gathered = [None] * P
            dist.all_gather_object(gathered, stores[dst])
            if rank == dst:
                inbox = gathered
        flat = np.full(nkt * ny * nzl * B, np.nan + 0j)
        for ad, val in inbox:
            flat[ad] = val
        by = flat[(((kx // B) * ny + yy) * nzl + zl) * B + kx % B]                 # what the inverse y-pass loads: (nkr, ny, nzl)
        back = sfft.irfft(sfft.ifft(by, axis=1), n=nx, axis=0)
        assert np.linalg.norm(back - xl) / np.linalg.norm(xl) < 1e-13
        # 2c. 2-D slab decomposition (csrc/fft_plan.cu::exec_dist2d): physical y-slabs, spectral kx blocks of kb = nx/(2P) wavenumbers
        #     plus one extra column (Nyquist on the last rank, zero padding elsewhere); the x pass stores each line's half spectrum
        #     block by block in destination-rank-major order [peer][kb + 1, ny_local]; equal blocks -> plain all-to-all, no pack
        nx2, ny2 = 16, 12
        x2 = np.asfortranarray(rng.standard_normal((nx2, ny2)))
        ref2 = sfft.rfftn(x2, axes=(1, 0))
        kb, nyl2 = nx2 // 2 // P, ny2 // P
        xl2 = ff.physical_slab_2d(x2, P, rank)
        a2 = sfft.rfft(xl2, axis=0)                                                # (nkr, nyl)
        send2 = np.zeros((P, kb + 1, nyl2), dtype=complex)
        for k in range(nx2 // 2 + 1):
            qd, kk = (P - 1, kb) if k == nx2 // 2 else (k // kb, k % kb)          # Pow2Params::row_seg_* / row_nyq
            send2[qd, kk, :] = a2[k, :]
        inbox2 = [None] * P
        for dst in range(P):
            gathered = [None] * P
            dist.all_gather_object(gathered, send2[dst])
            if rank == dst:
                inbox2 = gathered
        slab = np.zeros((kb + 1, ny2), dtype=complex, order="F")
        for s_ in range(P):
            slab[:, s_ * nyl2:(s_ + 1) * nyl2] = inbox2[s_]                        # rows of sender s land at y = s*nyl + y_local
        spec2 = sfft.fft(slab, axis=1)
        want2 = ff.spectral_slab_2d(ref2, P, rank)
        assert np.linalg.norm(spec2 - want2) / np.linalg.norm(ref2) < 1e-13
        if rank != P - 1:
            assert np.all(spec2[kb] == 0)                                          # padding column stays exactly zero
        everyone = [None] * P
        dist.all_gather_object(everyone, spec2)
        assert np.linalg.norm(ff.gather_spectral_2d(everyone, P) - ref2) / np.linalg.norm(ref2) < 1e-13
        kal = ff.getaliasedwavenumbers(nx2, nx2 // 2 + 1, 1 / 3)[1]               # kralias = iL:nkr
        lk = ff.local_kx_alias_2d(kal, nx2, P, rank)
        gmask = np.zeros(nx2 // 2 + 1, bool)
        gmask[kal[0] - 1:kal[1]] = True
        lmask = np.zeros(kb + 1, bool)
        lmask[lk[0] - 1:lk[1]] = True
        assert np.array_equal(lmask[:kb], gmask[rank * kb:(rank + 1) * kb]) and lmask[kb]
        # 3. alias ranges on the slab
        lal = ff.getaliasedwavenumbers(ny, ny // 2 + 1, 1 / 3)[0]
        loc = ff.local_alias_range(lal, ny, P, rank)
        mask = np.zeros(ny, bool)
        mask[lal[0] - 1:lal[1]] = True
        lo, hi = ff.slab_range(ny, P, rank)
        expect = mask[lo:hi]
        mine = np.zeros(nyl, bool)
        if loc:
            mine[loc[0] - 1:loc[1]] = True
        assert np.array_equal(mine, expect)
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_slab_logic_world_size_2_gloo():
    import torch.multiprocessing as mp
    import __graft_entry__ as ge
    ge.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def test_partition_helpers():
    import fourierflows_jl_b200 as ff
    assert ff.slab_range(1024, 8, 3) == (384, 512)
    assert ff.local_alias_range((342, 683), 1024, 8, 2) == (86, 128)
    assert ff.local_alias_range((342, 683), 1024, 8, 7) is None
    assert ff.local_alias_range(None, 1024, 8, 0) is None
    with pytest.raises(ff.FFBError):
        ff.slab_range(10, 4, 0)
    # 0.94 GB per GPU per 3-D FFT at 1024^3 Float64 on 8 GPUs
    assert abs(ff.exchange_bytes_per_rank((1024, 1024, 1024), 8, 8) / 1e9 - 0.94) < 0.01
