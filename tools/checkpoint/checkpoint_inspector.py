#!/usr/bin/env python
"""Look inside a distributed checkpoint without building a model (reference ``tools/checkpoint/checkpoint_inspector.py``).

    python tools/checkpoint/checkpoint_inspector.py inspect  /ckpt/iter_0001000            # keys, global shapes, dtypes, bytes
    python tools/checkpoint/checkpoint_inspector.py args     /ckpt/iter_0001000            # the training arguments saved with it
    python tools/checkpoint/checkpoint_inspector.py diff     /ckpt/a /ckpt/b [--values]    # key / shape differences, optionally max |Δ|
    python tools/checkpoint/checkpoint_inspector.py rename   /ckpt/a /ckpt/out --sub 'decoder\\.' 'encoder.'   # rewrite tensor keys
"""
import argparse
import os
import re
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def _reader(path):
    from torch.distributed.checkpoint import FileSystemReader

    return FileSystemReader(path)


def read_metadata(path):
    """→ ``{key: (global shape, dtype)}`` for tensors, ``{key: None}`` for pickled objects."""
    from torch.distributed.checkpoint.metadata import BytesStorageMetadata

    md = _reader(path).read_metadata()
    out = {}
    for k, v in md.state_dict_metadata.items():
        out[k] = None if isinstance(v, BytesStorageMetadata) else (tuple(v.size), v.properties.dtype)
    return out


def load_full(path, keys=None):
    """Materialise (a subset of) the tensors, unsharded, on the CPU in one process."""
    import torch.distributed.checkpoint as dcp

    meta = read_metadata(path)
    sd = {k: torch.empty(s, dtype=d) for k, v in meta.items() if v is not None and (keys is None or k in keys) for s, d in [v]}
    dcp.load(sd, storage_reader=_reader(path), no_dist=True)
    return sd


def cmd_inspect(a):
    meta = read_metadata(a.path)
    total = 0
    for k in sorted(meta):
        if a.filter and not re.search(a.filter, k):
            continue
        if meta[k] is None:
            print(f"{k:80s} <object>")
            continue
        shape, dtype = meta[k]
        n = 1
        for s in shape:
            n *= s
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        total += nbytes
        print(f"{k:80s} {str(shape):24s} {str(dtype):16s} {nbytes / 2**20:10.2f} MiB")
    print(f"{len(meta)} entries, {total / 2**30:.3f} GiB of tensors")
    return meta


def cmd_args(a):
    p = os.path.join(a.path, "common.pt")
    if not os.path.exists(p):
        raise SystemExit(f"no common.pt under {a.path}")
    common = torch.load(p, map_location="cpu", weights_only=False)
    args = common.get("args")
    for k, v in sorted(vars(args).items() if args is not None else []):
        print(f"{k} = {v}")
    print(f"iteration = {common.get('iteration')}")
    return common


def cmd_diff(a):
    ma, mb = read_metadata(a.path), read_metadata(a.other)
    only_a, only_b = sorted(set(ma) - set(mb)), sorted(set(mb) - set(ma))
    for k in only_a:
        print(f"- {k}")
    for k in only_b:
        print(f"+ {k}")
    changed = [k for k in sorted(set(ma) & set(mb)) if ma[k] != mb[k]]
    for k in changed:
        print(f"~ {k}: {ma[k]} -> {mb[k]}")
    worst = {}
    if a.values:
        same = [k for k in set(ma) & set(mb) if ma[k] == mb[k] and ma[k] is not None]
        ta, tb = load_full(a.path, set(same)), load_full(a.other, set(same))
        for k in sorted(same):
            d = (ta[k].float() - tb[k].float()).abs().max().item() if ta[k].numel() else 0.0
            if d > a.atol:
                worst[k] = d
                print(f"! {k}: max |delta| = {d:.3e}")
    print(f"{len(only_a)} only in A, {len(only_b)} only in B, {len(changed)} with different shape/dtype" + (f", {len(worst)} with different values" if a.values else ""))
    return only_a, only_b, changed, worst


def cmd_rename(a):
    import torch.distributed.checkpoint as dcp
    from torch.distributed.checkpoint import FileSystemWriter

    sd = load_full(a.path)
    out = {}
    for k, v in sd.items():
        nk = k
        for pat, rep in a.sub or []:
            nk = re.sub(pat, rep, nk)
        if nk in out:
            raise SystemExit(f"rename maps two keys onto {nk}")
        out[nk] = v
    os.makedirs(a.other, exist_ok=True)
    dcp.save(out, storage_writer=FileSystemWriter(a.other), no_dist=True)
    for f in ("common.pt", "metadata.json"):
        if os.path.exists(os.path.join(a.path, f)):
            import shutil

            shutil.copy(os.path.join(a.path, f), os.path.join(a.other, f))
    print(f"wrote {len(out)} tensors to {a.other}")
    return out


def main(argv=None):
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    p = sub.add_parser("inspect"); p.add_argument("path"); p.add_argument("--filter", default=None); p.set_defaults(fn=cmd_inspect)
    p = sub.add_parser("args"); p.add_argument("path"); p.set_defaults(fn=cmd_args)
    p = sub.add_parser("diff"); p.add_argument("path"); p.add_argument("other"); p.add_argument("--values", action="store_true"); p.add_argument("--atol", type=float, default=0.0); p.set_defaults(fn=cmd_diff)
    p = sub.add_parser("rename"); p.add_argument("path"); p.add_argument("other"); p.add_argument("--sub", nargs=2, action="append", metavar=("PATTERN", "REPL")); p.set_defaults(fn=cmd_rename)
    a = ap.parse_args(argv)
    return a.fn(a)


if __name__ == "__main__":
    main()
