"""Drop-in for the hot-path surface of subphaser/Seqs.py (reference v1.2.7): `map_kmer3` (:74-119) with
the same signature, output file format and log lines.  The per-base Python loop of `map_kmer_each4`
(:209-237) becomes the K9 kernel: every position of the 2-bit packed chromosome (already resident in
HBM from the counting step) is looked up in the specific-k-mer table and counted into one row per
(10-kb bin, 10-Mb chunk) pair, which reproduces the reference's line structure — including the
duplicate-coordinate lines at chunk borders (:131-137, :229-236) — byte for byte when the reference
runs with the ordered `method='map'`.

Multi-record inputs (custom features / LTRs, `chunk=False`) go through the same kernel with a record index.

`split_genomes` (:27-71, FASTA re-writing with BioPython in the reference) runs on the device as well: the records of
a genome file are found by `spk_fasta_record_starts`, every wanted chromosome is packed by K1 straight from the device
copy of the file and registered for the counting / mapping steps, and the per-chromosome FASTA files are written from
verbatim byte ranges of the input whenever `spk_fasta_wrap_check` finds them already in 60-column layout.
"""
import copy
import logging
import os
import re
import sys
from collections import OrderedDict

import numpy as np

from . import _registry, engine, kmer_codec

logger = logging.getLogger("subphaser_b200")


def _first_id(path):
    with open(path, "rb") as f:
        head = f.read(65536)
    if head[:2] == b"\x1f\x8b":
        import gzip
        with gzip.open(path, "rb") as f:
            head = f.read(65536)
    line = head.split(b"\n", 1)[0].decode(errors="replace")
    if not line.startswith(">"):
        raise ValueError("{} is not a FASTA file".format(path))
    t = line[1:].split()
    return t[0] if t else ""


def _packed(chromfile):
    seq = _registry.get_seq(chromfile)
    if seq is None:
        d, n = engine.to_device_bytes(engine.read_fasta_bytes(chromfile))
        seq = engine.pack_fasta(d, n, name=os.path.basename(chromfile))
        _registry.put_seq(chromfile, seq)
    return seq


def _sig_table(d_kmers, k, sg_names):
    """d_kmers (KmerSGMap from Cluster.output_kmers, or a plain dict kmer -> SG) -> engine.SigTable."""
    import torch
    from .Cluster import KmerSGMap
    names = list(sg_names)
    if isinstance(d_kmers, KmerSGMap):
        keys = d_kmers.keys_arr
        remap = np.array([names.index(s) for s in d_kmers.sg_names], dtype=np.uint8)
        vals = remap[d_kmers.sg_idx] if len(keys) else np.zeros(0, np.uint8)
        k = d_kmers.k if k is None else k
    else:
        strs = [s for s in d_kmers.keys() if isinstance(s, str)]
        if k is None and strs:
            k = len(strs[0])
        strs = [s for s in strs if len(s) == k and s == s.upper()]
        keys, valid = kmer_codec.strs_to_keys(strs, k)
        vals = np.array([names.index(d_kmers[s]) for s in strs], dtype=np.uint8)
        keys, vals = keys[valid], vals[valid]
        canon = kmer_codec.canonical_keys(keys, k) if len(keys) else keys
        # a forward-only entry (no reverse complement in the dict) cannot be expressed canonically
        if len(keys):
            fwd = set(keys.tolist())
            rc = kmer_codec.revcomp_keys(keys, k)
            if any(int(r) not in fwd for r in rc):
                raise NotImplementedError("d_kmers must hold every k-mer together with its reverse complement")
        order = np.argsort(canon, kind="stable")
        canon, vals = canon[order], vals[order]
        first = np.ones(len(canon), bool)
        first[1:] = canon[1:] != canon[:-1]
        keys, vals = canon[first], vals[first]
    dev = engine._dev()
    dk = torch.from_numpy(np.ascontiguousarray(keys).view(np.int64).copy()).to(dev)
    dv = torch.from_numpy(np.ascontiguousarray(vals)).to(dev)
    return engine.SigTable(dk, dv, k, S=len(names)), k


def _lines_of(seq_len, k, bin_size, chunk_size):
    """Decode line ids -> (bin start, end) exactly as map_kmer_each4 prints them (Seqs.py:228-236)."""
    n = engine._lib.load().spk_map_num_lines(seq_len, k, bin_size, chunk_size)
    return n


def map_kmer3(chromfiles, d_kmers, fout=sys.stdout, k=None, window_size=10e6,
              bin_size=10000, sg_names=[],
              ncpu="autodetect", method="map", log=True, chunk=True, chunksize=None):
    if k is None:
        for key in d_kmers.keys():
            k = len(key)
            break
    engine.require_cuda()
    S = len(sg_names)
    sig, k = _sig_table(d_kmers, k, sg_names)
    bin_size = int(bin_size)
    W = int(window_size) if chunk else 0
    fout.write("\t".join(["#chrom", "start", "end"] + list(sg_names)) + "\n")
    i = 0
    mapped_num, mapped_seqs = 0, 0
    all_lines = []
    lib = engine._lib.load()
    for chromfile in chromfiles:
        seq = _packed(chromfile)
        if seq.n_records > 1:
            # every record is mapped on its own coordinates (Seqs.py:121-153); one kernel pass over the file
            recs = [(rid, len(rseq)) for rid, _, rseq in _iter_fasta(chromfile)]
            if sum(L + 1 for _, L in recs) - 1 != seq.n_bases:
                raise ValueError("cannot index the records of {} (text before the first header?)".format(chromfile))
            counts_d, _ = engine.map_bins(seq, sig, S, bin_size, W, record_lengths=[L for _, L in recs])
        else:
            recs = [(_first_id(chromfile), seq.n_bases)]
            counts_d, _ = engine.map_bins(seq, sig, S, bin_size, W)
        counts_all = counts_d.cpu().numpy().view(np.uint32)
        row0 = 0
        for cid, L in recs:
            nl = lib.spk_map_num_lines(L, k, bin_size, W)
            counts = counts_all[row0:row0 + nl]
            row0 += nl
            if chunk:
                logger.info("Chunking chromsome {}: {:,} bp".format(cid, L))
            nhits = int(counts.sum(dtype=np.int64))
            nz = np.nonzero(counts.any(axis=1))[0]
            # line id -> (bin, chunk): walk the (bin, chunk) boundaries in position order
            if len(nz):
                lid = nz.astype(np.int64)
                if W:
                    # line = pos//bin + (pos+k-1)//W is monotone in pos; recover bin by searching the
                    # first position of every line: bins and chunk starts are the breakpoints
                    brk = np.unique(np.concatenate([
                        np.arange(0, L, bin_size, dtype=np.int64),
                        np.maximum(np.arange(0, L + k, W, dtype=np.int64) - (k - 1), 0)]))
                    brk = brk[brk < L]
                    first_line = brk // bin_size + (brk + k - 1) // W
                    pos = brk[np.searchsorted(first_line, lid)]
                    bins = pos // bin_size
                    chk = (pos + k - 1) // W
                    size = np.minimum((chk + 1) * W, L)
                else:
                    bins = lid
                    size = np.full(len(lid), L, dtype=np.int64)
                starts = bins * bin_size
                ends = np.minimum(starts + bin_size, size)
                rows = counts[nz]
                text = "".join(
                    "{}\t{}\t{}\t{}\n".format(cid, s, e, "\t".join(map(str, r)))
                    for s, e, r in zip(starts.tolist(), ends.tolist(), rows.tolist()))
                fout.write(text)
                all_lines.append((cid, starts, ends, rows.astype(np.int64)))
            if log:
                logger.info("Mapped {} kmers to chromsome {}".format(nhits, cid))
            n_chunks = max(1, -(-L // W)) if W else 1
            i += n_chunks
            mapped_num += nhits
            if nhits > 0:
                # the reference counts the chunks ("sequences") that contain a hit (Seqs.py:110-111): a line id encodes
                # its chunk, so the chunks with hits are the distinct chunk ids of the non-empty lines
                mapped_seqs += int(len(np.unique(chk))) if (W and len(nz)) else 1
    logger.info("Processed {} sequences".format(i))
    mapped_cat, total = sig.n_mapped(), len(d_kmers)
    try:
        logger.info("{} ({:.2%}) sequences contain subgenome-specific kmers".format(mapped_seqs, mapped_seqs / i))
        logger.info("{:.2%} of {} subgenome-specific kmers are mapped".format(mapped_cat / total, total // 2))
    except ZeroDivisionError:
        logger.warning("None sequences, please check.")
    name = getattr(fout, "name", None)
    if isinstance(name, str):
        fout.flush()
        if os.path.exists(name):
            _registry.put_bins(name, all_lines)


def _target_maps(targets, d_targets, sep):
    """The two id maps of Seqs.py:29-46.  d_targets: id as it occurs in the genome files (possibly with the genome's
    prefix) -> output id; d_targets2: entry as written in the config -> output id.  A config entry `new|old` renames."""
    d2 = OrderedDict()
    if not d_targets:
        d_targets = OrderedDict()
        todo = list(targets)
    else:
        todo = list(set(targets) - set(d_targets))
        if not todo:
            return d_targets, copy.deepcopy(d_targets)
    for entry in todo:
        new_id, _, old_id = entry.partition(sep)
        d_targets[old_id if _ else new_id] = new_id
        d2[entry] = new_id
    return d_targets, d2


def _read_genome_bytes(path):
    """Genome file -> uint8 array in pinned host memory (gzip members inflated by zlib in a worker thread while the
    previous file is on the GPU: the reference's `zcat` leg, Jellyfish.py:696)."""
    import torch
    with open(path, "rb") as f:
        magic = f.read(2)
    if magic == b"\x1f\x8b":
        import zlib
        pieces, size = [], 0
        with open(path, "rb") as f:
            dec = zlib.decompressobj(wbits=31)
            while True:
                raw = f.read(1 << 24)
                if not raw:
                    break
                while raw:
                    out = dec.decompress(raw)
                    if out:
                        pieces.append(out)
                        size += len(out)
                    if dec.eof:                      # concatenated gzip members (bgzip)
                        raw = dec.unused_data
                        dec = zlib.decompressobj(wbits=31)
                    else:
                        raw = b""
        host = torch.empty(size + 16, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
        view = host.numpy()
        at = 0
        for piece in pieces:
            view[at:at + len(piece)] = np.frombuffer(piece, np.uint8)
            at += len(piece)
        return host, size
    size = os.path.getsize(path)
    host = torch.empty(size + 16, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
    with open(path, "rb") as f:
        f.readinto(memoryview(host.numpy())[:size])
    return host, size


def _write_chrom_file(path, title, body, verbatim, wrap=60):
    """`>title` + the sequence in lines of 60 (what SeqIO.write produces).  verbatim: `body` (uint8 view of the input
    file) already has that layout."""
    with open(path, "wb") as f:
        f.write(b">" + title.encode() + b"\n")
        if verbatim:
            f.write(memoryview(body))
            if len(body) and body[-1] != 10:
                f.write(b"\n")
            return
        seq = body[(body != 10) & (body != 13) & (body != 32)]
        n = len(seq)
        full = n // wrap
        if full:
            lines = np.empty((full, wrap + 1), np.uint8)
            lines[:, :wrap] = seq[:full * wrap].reshape(full, wrap)
            lines[:, wrap] = 10
            f.write(memoryview(lines.reshape(-1)))
        if n % wrap:
            f.write(memoryview(seq[full * wrap:]))
            f.write(b"\n")


def split_genomes(genomes, prefixes, targets, outdir, d_targets=None, sep="|"):
    """Seqs.py:27-71 with the genome on the GPU instead of BioPython on the host.  Per genome file: the bytes go to the
    device once (pinned buffer; gz inflated by zlib), `spk_fasta_record_starts` finds the records, the ids are resolved on
    the host (a few bytes per record), and every wanted record is (a) packed by K1 straight from the device buffer — the
    2-bit sequence is registered under the chromosome's file path, so the count and map steps never read the file —
    and (b) written as `<outdir><id>.fasta` by a writer thread, as a verbatim byte range of the input when
    `spk_fasta_wrap_check` finds it already in 60-column layout.  Returns (files, labels, d_targets2, d_size) as the
    reference does; `d_size[id]` = number of sequence characters."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    from . import _lib
    engine.require_cuda()
    d_targets, d_targets2 = _target_maps(targets, d_targets, sep)
    outfas, labels, d_size, got_ids = [], [], {}, set()
    pending = []
    dev = engine._dev()
    with ThreadPoolExecutor(max_workers=4) as writers, ThreadPoolExecutor(max_workers=1) as reader:
        nxt = reader.submit(_read_genome_bytes, genomes[0]) if genomes else None
        for gi, (genome, prefix) in enumerate(zip(genomes, prefixes)):
            host, size = nxt.result()
            nxt = reader.submit(_read_genome_bytes, genomes[gi + 1]) if gi + 1 < len(genomes) else None
            d_file = torch.empty(size + 16, dtype=torch.uint8, device=dev)
            d_file[:size].copy_(host[:size], non_blocking=True)
            view = host.numpy()[:size]
            # ---- record starts (device) ----
            cap = 1 << 16
            while True:
                d_pos = torch.empty(cap, dtype=torch.int64, device=dev)
                d_cnt = torch.zeros(1, dtype=torch.int64, device=dev)
                _lib.call("spk_fasta_record_starts", engine._p(d_file), size, engine._p(d_pos), cap, engine._p(d_cnt),
                          engine._stream())
                n_rec = int(d_cnt.item())
                if n_rec <= cap:
                    break
                cap = n_rec
            starts = np.sort(d_pos[:n_rec].cpu().numpy()).tolist() + [size]
            for r in range(n_rec):
                a, b = starts[r], starts[r + 1]
                eol = a + int(np.argmax(view[a:min(b, a + 65536)] == 10)) if (view[a:min(b, a + 65536)] == 10).any() else b
                title = bytes(view[a + 1:eol]).decode(errors="replace").rstrip("\r")
                old_id = title.split(None, 1)[0] if title.split() else ""
                rid = prefix + old_id
                if d_targets:
                    if rid in d_targets:
                        pass
                    elif old_id in d_targets:
                        rid = old_id
                    else:
                        continue
                got_ids.add(rid)
                new_id = d_targets[rid]                 # (no targets at all: KeyError, as in the reference)
                # FastaWriter: the old title is kept behind a changed id
                out_title = title if (title.split(None, 1)[:1] == [new_id]) else ("{} {}".format(new_id, title) if title else new_id)
                body_a = min(eol + 1, b)
                # ---- K1 on the record's bytes (device) ----
                nb = b - a
                d_rec = torch.empty(nb + 16, dtype=torch.uint8, device=dev)
                d_rec[:nb].copy_(d_file[a:b])
                seq = engine.pack_fasta(d_rec, nb, name=new_id)
                del d_rec
                d_chk = torch.zeros(2, dtype=torch.int64, device=dev)
                _lib.call("spk_fasta_wrap_check", engine._p(d_file), body_a, b, 60, engine._p(d_chk), engine._stream())
                flags = int(d_chk[0].item())
                outfa = "{}{}.fasta".format(outdir, new_id)
                fut = writers.submit(_write_chrom_file, outfa, out_title, view[body_a:b], flags == 0)
                pending.append((fut, outfa, seq))
                outfas.append(outfa)
                labels.append(new_id)
                d_size[new_id] = seq.n_bases
            torch.cuda.current_stream().synchronize()
            for fut, outfa, seq in pending:
                fut.result()
                _registry.put_seq(outfa, seq)          # (after the file exists: the registry keys on its signature)
            pending = []
            del d_file, host
    ungot_ids = set(d_targets) - got_ids
    if ungot_ids:
        logger.error("Chromosomes {} are not found in sequences files".format(ungot_ids))
    return outfas, labels, d_targets2, d_size


def _iter_fasta(path):
    opener = open
    with open(path, "rb") as f:
        if f.read(2) == b"\x1f\x8b":
            import gzip
            opener = gzip.open
    rid, desc, chunks = None, "", []
    with opener(path, "rt") as f:
        for line in f:
            line = line.rstrip("\r\n")
            if line.startswith(">"):
                if rid is not None:
                    yield rid, desc, "".join(chunks)
                t = line[1:].split(None, 1)
                rid = t[0] if t else ""
                desc = t[1] if len(t) > 1 else ""
                chunks = []
            elif rid is not None:
                chunks.append(line)
    if rid is not None:
        yield rid, desc, "".join(chunks)
