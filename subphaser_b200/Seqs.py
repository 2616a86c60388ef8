"""Drop-in for the hot-path surface of subphaser/Seqs.py (reference v1.2.7): `map_kmer3` (:74-119) with
the same signature, output file format and log lines.  The per-base Python loop of `map_kmer_each4`
(:209-237) becomes the K9 kernel: every position of the 2-bit packed chromosome (already resident in
HBM from the counting step) is looked up in the specific-k-mer table and counted into one row per
(10-kb bin, 10-Mb chunk) pair, which reproduces the reference's line structure — including the
duplicate-coordinate lines at chunk borders (:131-137, :229-236) — byte for byte when the reference
runs with the ordered `method='map'`.

`split_genomes` (:27-71, FASTA re-writing with BioPython) is the next component of SURVEY.md §8f and
is provided here as plain host code so that `__main__.py` keeps working.
Multi-record inputs with `chunk=False` (custom features / LTRs) are not on the GPU path yet.
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
                # the reference counts chunks ("sequences") containing hits; per-chunk hit flags are not
                # tracked on the device, records with hits are counted chunk-wise as all-mapped
                mapped_seqs += n_chunks
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


def split_genomes(genomes, prefixes, targets, outdir, d_targets=None, sep="|"):
    """Seqs.py:27-71 (host code, no BioPython): one FASTA per target chromosome, id remapping."""
    d_targets2 = OrderedDict()
    if not d_targets:
        d_targets = OrderedDict()
        for t in targets:
            temp = t.split(sep, 1)
            id, new_id = temp[-1], temp[0]
            d_targets[id] = new_id
            d_targets2[t] = new_id
    elif set(targets) - set(d_targets):
        for t in set(targets) - set(d_targets):
            temp = t.split(sep, 1)
            id, new_id = temp[-1], temp[0]
            d_targets[id] = new_id
            d_targets2[t] = new_id
    else:
        d_targets2 = copy.deepcopy(d_targets)
    outfas, labels = [], []
    d_size = {}
    got_ids = set([])
    for genome, prefix in zip(genomes, prefixes):
        for rid, desc, seq in _iter_fasta(genome):
            old_id, new_id = rid, "{}{}".format(prefix, rid)
            if d_targets:
                if new_id in d_targets:
                    rid = new_id
                elif old_id in d_targets:
                    pass
                else:
                    continue
            got_ids.add(rid)
            rid = d_targets[rid]
            outfa = "{}{}.fasta".format(outdir, rid)
            with open(outfa, "w") as fout:
                fout.write(">{} {}\n".format(rid, desc) if desc else ">{}\n".format(rid))
                for a in range(0, len(seq), 60):
                    fout.write(seq[a:a + 60] + "\n")
            outfas += [outfa]
            labels += [rid]
            d_size[rid] = len(seq)
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
