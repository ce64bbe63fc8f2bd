#!/usr/bin/env python
"""chunk pairs per second THROUGH THE TCP PATH -- a scripted master (the reference's wire protocol,
analysis/WorkQueue.cc:71-166) hands GridSearch-style 24x24-chunk blocks of a synthetic genome pair to the B200 slave
executable over loopback, `per_exchange` blocks per connection like the reference master's 2 x threads, and counts the
records it gets back; next to it the same blocks through the C ABI in process.  Run on the GPU box:
    python tools/slave_throughput.py [megabases] [blocks per exchange]   -> one JSON line."""
import json
import os
import socket
import struct
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import satsuma2_b200 as sx  # noqa: E402
from satsuma2_b200 import build as sxbuild, synth  # noqa: E402


def recv_all(c, n):
    buf = bytearray()
    while len(buf) < n:
        part = c.recv(min(n - len(buf), 1 << 20))
        if not part:
            raise IOError("short read")
        buf += part
    return buf


def main():
    mb = float(sys.argv[1]) if len(sys.argv) > 1 else 24.0
    per = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    L = int(mb * 1e6)
    tgt, qry = synth.genome_pair(L, seed=11)
    tmp = tempfile.mkdtemp()
    qf, tf = os.path.join(tmp, "q.fa"), os.path.join(tmp, "t.fa")
    for path, name, seq in ((qf, "q", qry), (tf, "t", tgt)):
        with open(path, "w") as f:
            f.write(">" + name + "\n")
            s = seq.tobytes().decode()
            for i in range(0, len(s), 100):
                f.write(s[i:i + 100] + "\n")
    to, tl, ts = synth.chunk_sequence(tgt, 4096, 1024)
    qo, ql, qs = synth.chunk_sequence(qry, 4096, 0)
    blocks = sx.make_blocks(synth.diagonal_blocks(len(tl), len(ql), 3072, 4096, pixel=24))
    n_pairs = int(((blocks["target_to"] - blocks["target_from"] + 1).astype(np.int64) *
                   (blocks["query_to"] - blocks["query_from"] + 1)).sum())
    # ---- in process, for comparison
    with sx.XCorrEngine(target_total=float(L)) as eng:
        eng.set_targets(sx.ChunkSet(tgt, to, tl, ts, np.zeros(len(tl), np.int32), [L]))
        eng.set_queries(sx.ChunkSet(qry, qo, ql, qs, np.zeros(len(ql), np.int32), [L]))
        eng.align_blocks(blocks)
        eng.invalidate_spectra()
        t0 = time.perf_counter()
        recs = eng.align_blocks(blocks)
        t_inproc = time.perf_counter() - t0
    # ---- through the slave
    exe = [e for e in sxbuild.build_host() if e.endswith("HomologyByXCorrSlave")][0]
    srv = socket.socket()
    srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
    srv.bind(("127.0.0.1", 0))
    srv.listen(64)
    port = srv.getsockname()[1]
    state = {"sent": 0, "records": 0, "connections": 0, "first": None, "last": None, "done": False, "last_data": None}
    raw = blocks.tobytes()

    def master():
        while not state["done"]:
            c, _ = srv.accept()
            c.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
            state["connections"] += 1
            sid, n = struct.unpack("<II", bytes(recv_all(c, 8)))
            now = time.perf_counter()
            if n:
                recv_all(c, 72 * n)
                state["records"] += n
                state["last_data"] = time.perf_counter()
            if state["sent"] < len(blocks):
                if state["first"] is None:
                    state["first"] = now
                k = min(per, len(blocks) - state["sent"])
                c.sendall(struct.pack("<i", k) + raw[28 * state["sent"]:28 * (state["sent"] + k)])
                state["sent"] += k
                state["last_data"] = time.perf_counter()
            elif state["records"] >= len(recs) or time.perf_counter() - state["last_data"] > 5.0:
                state["last"] = state["last_data"]
                c.sendall(struct.pack("<i", -1))
                state["done"] = True
            else:
                c.sendall(struct.pack("<i", 0))
            c.close()

    th = threading.Thread(target=master, daemon=True)
    th.start()
    p = subprocess.run([exe, "-master", "127.0.0.1", "-port", str(port), "-sid", "1", "-q", qf, "-t", tf, "-device", "0"],
                       capture_output=True, text=True, timeout=900)
    th.join(timeout=30)
    srv.close()
    t_tcp = state["last"] - state["first"]
    tail = [ln for ln in p.stdout.splitlines() if ln.startswith("exchanges")]
    print(json.dumps({"what": f"{mb:g} Mb x {mb:g} Mb synthetic genome pair, {len(blocks)} blocks of 24x24 chunks = {n_pairs} chunk "
                              f"pairs; scripted master on loopback TCP handing out {per} blocks per connection (python, "
                              "one thread) vs sx_align_blocks in process (cold spectra)",
                      "chunk_pairs": n_pairs, "records_in_process": int(len(recs)), "records_over_tcp": state["records"],
                      "in_process_s": t_inproc, "in_process_pairs_per_s": n_pairs / t_inproc,
                      "tcp_s": t_tcp, "tcp_pairs_per_s": n_pairs / t_tcp, "connections": state["connections"],
                      "slave_says": tail[0] if tail else "", "slave_rc": p.returncode}))


if __name__ == "__main__":
    main()
