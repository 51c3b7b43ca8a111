"""FASTA parse/write as Biopython does it for the calls in the reference (60-column wrap on write)."""
from .Seq import Seq
from .SeqRecord import SeqRecord


def parse(handle, fmt):
    assert fmt == "fasta"
    close = False
    if isinstance(handle, str):
        handle = open(handle)
        close = True
    name = None
    chunks = []
    for line in handle:
        line = line.rstrip("\r\n")
        if line.startswith(">"):
            if name is not None:
                yield _mk(name, chunks)
            name = line[1:]
            chunks = []
        elif name is not None:
            chunks.append(line.strip())
    if name is not None:
        yield _mk(name, chunks)
    if close:
        handle.close()


def _mk(title, chunks):
    parts = title.split(None, 1)
    rid = parts[0] if parts else ""
    return SeqRecord(Seq("".join(chunks)), id=rid, name=rid, description=title)


def write(records, handle, fmt):
    assert fmt == "fasta"
    close = False
    if isinstance(handle, str):
        handle = open(handle, "w")
        close = True
    n = 0
    for r in records:
        title = r.id if not r.description or r.description == r.id else (r.id + " " + r.description if not r.description.startswith(r.id) else r.description)
        handle.write(">%s\n" % title)
        s = str(r.seq)
        for i in range(0, len(s), 60):
            handle.write(s[i:i + 60] + "\n")
        n += 1
    if close:
        handle.close()
    return n
