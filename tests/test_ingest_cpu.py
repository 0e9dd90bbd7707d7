"""CPU tests of the native corpus ingest (tmvb_read_docfile): readcorp's docfile semantics
(Corpus.jl:277-296) + the flattening of update_buffer! (modelutils.jl:371-380), against a pure-Python restatement."""
import os

import numpy as np
import pytest


def _py_readcorp(path, delim=",", counts=False, readers=False, ratings=False):
    """Corpus.jl:288-295 restated: blocks of L lines, parse(Int, .) per field, Document defaults, check_doc."""
    if ratings and not readers:
        ratings = False
    with open(path, "rb") as f:
        text = f.read().decode()
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    lines = [l[:-1] if l.endswith("\r") else l for l in lines]
    L = 1 + counts + readers + ratings
    docs = []
    for d in range(0, len(lines), L):
        block = [[int(p) for p in line.split(delim)] for line in lines[d:d + L]]
        names = [n for n, on in zip(("terms", "counts", "readers", "ratings"), (True, counts, readers, ratings)) if on]
        doc = dict(zip(names, block))
        doc.setdefault("counts", [1] * len(doc["terms"]))
        doc.setdefault("readers", [])
        doc.setdefault("ratings", [1] * len(doc["readers"]))
        assert all(x > 0 for k in doc for x in doc[k]) and len(doc["terms"]) == len(doc["counts"]) and len(doc["readers"]) == len(doc["ratings"])
        docs.append(doc)
    return docs


def _write(path, docs, delim, counts, readers, ratings, crlf=False, blanks=False, final_newline=True):
    fmt = (lambda v: (" %d " % v)) if blanks else str
    rows = []
    for d in docs:
        rows.append(delim.join(fmt(v) for v in d["terms"]))
        if counts:
            rows.append(delim.join(fmt(v) for v in d["counts"]))
        if readers:
            rows.append(delim.join(fmt(v) for v in d["readers"]))
        if ratings:
            rows.append(delim.join(fmt(v) for v in d["ratings"]))
    eol = "\r\n" if crlf else "\n"
    with open(path, "wb") as f:
        f.write((eol.join(rows) + (eol if final_newline else "")).encode())


def _random_docs(rng, M, V, U):
    docs = []
    for _ in range(M):
        n, r = int(rng.integers(1, 40)), int(rng.integers(1, 9))
        docs.append(dict(terms=rng.choice(V, n, replace=False) + 1, counts=rng.integers(1, 30, n),
                         readers=rng.choice(U, r, replace=False) + 1, ratings=rng.integers(1, 6, r)))
    return docs


@pytest.mark.parametrize("counts,readers,ratings", [(False, False, False), (True, False, False), (True, True, False), (True, True, True),
                                                    (False, True, True)])
@pytest.mark.parametrize("style", ["plain", "crlf_blanks_nofinal", "tabs"])
def test_native_readcorp_matches_python_restatement(tm, tmp_path, counts, readers, ratings, style):
    rng = np.random.default_rng(3)
    docs = _random_docs(rng, 700, 900, 60)      # > 256 documents per thread: the multi-threaded path
    p = str(tmp_path / "docs.txt")
    delim = "\t" if style == "tabs" else ","
    _write(p, docs, delim, counts, readers, ratings, crlf=style.startswith("crlf"), blanks=style.startswith("crlf"),
           final_newline=not style.endswith("nofinal"))
    ref = _py_readcorp(p, delim, counts, readers, ratings)
    f = tm.readcorp(docfile=p, delim=delim, counts=counts, readers=readers, ratings=ratings).flat()
    assert f.M == len(ref)
    np.testing.assert_array_equal(np.diff(f.N_cumsum), [len(d["terms"]) for d in ref])
    np.testing.assert_array_equal(f.terms, np.concatenate([d["terms"] for d in ref]) - 1)          # 0-based, modelutils.jl:371
    np.testing.assert_array_equal(f.counts, np.concatenate([d["counts"] for d in ref]))
    np.testing.assert_array_equal(np.diff(f.R_cumsum), [len(d["readers"]) for d in ref])
    if readers:
        np.testing.assert_array_equal(f.readers, np.concatenate([d["readers"] for d in ref]) - 1)
        np.testing.assert_array_equal(f.ratings, np.concatenate([d["ratings"] for d in ref]))
    assert f.V == max(int(d["terms"].max()) for d in docs) and f.U == (max(int(d["readers"].max()) for d in docs) if readers else 0)


def test_partial_last_block_takes_defaults(tm, tmp_path):
    """Iterators.partition yields a shorter last block; zip drops the missing keywords (Corpus.jl:288-291)."""
    p = str(tmp_path / "d.txt")
    open(p, "w").write("1,2,3\n4,5,6\n7,8\n")
    f = tm.readcorp(docfile=p, counts=True).flat()
    assert f.M == 2
    np.testing.assert_array_equal(f.terms, [0, 1, 2, 6, 7])
    np.testing.assert_array_equal(f.counts, [4, 5, 6, 1, 1])


@pytest.mark.parametrize("body,kw,doc,line", [
    ("1,2\n\n3\n", dict(), 2, 2),                      # empty line: parse(Int, "") throws
    ("1,2\n3,x\n", dict(), 2, 2),                      # not an integer
    ("1,2\n1,1\n3,0\n1,1\n", dict(counts=True), 2, 3),   # check_doc: terms must be positive
    ("1,2\n1\n", dict(counts=True), 1, 1),             # terms and counts of different length
    ("1,2\n1,1\n5\n3,4\n1,1\n5,6\n", dict(counts=True, readers=True), 2, 4),   # fine so far ...
])
def test_failing_document_raises_the_reference_message(tm, tmp_path, body, kw, doc, line):
    p = str(tmp_path / "d.txt")
    open(p, "w").write(body)
    if "5,6" in body:                                   # ... a well-formed file must load
        assert tm.readcorp(docfile=p, **kw).flat().M == 2
        return
    with pytest.raises(tm.CorpusError, match=r"document %d beginning on line %d failed to load\." % (doc, line)):
        tm.readcorp(docfile=p, **kw)


def test_vocab_and_user_files_size_the_corpus(tm, tmp_path):
    p, v, u = str(tmp_path / "d.txt"), str(tmp_path / "v.txt"), str(tmp_path / "u.txt")
    open(p, "w").write("1,3\n2,2\n2\n")
    open(v, "w").write("".join("%d\tw%d\n" % (i, i) for i in range(1, 6)))
    open(u, "w").write("1\ta\n2\tb\n3\tc\n")
    c = tm.readcorp(docfile=p, vocabfile=v, userfile=u, counts=True, readers=True)
    assert c.size() == (1, 5, 3)
    open(v, "w").write("1\tw1\n")
    with pytest.raises(tm.CorpusError, match="term keys not found"):
        tm.readcorp(docfile=p, vocabfile=v, counts=True, readers=True)


@pytest.mark.skipif(not os.path.exists("/root/reference/datasets/nsf/nsfdocs.txt"), reason="reference datasets not mounted")
def test_real_nsf_and_citeulike_files(tm):
    """readcorp(:nsf) / readcorp(:citeu) (Corpus.jl:337-352) through the native parser == tools/pack_corpus.py's output."""
    root = "/root/reference/datasets"
    c = tm.readcorp(docfile=root + "/nsf/nsfdocs.txt", vocabfile=root + "/nsf/nsfvocab.txt", counts=True)
    assert c.size() == (128804, 25319, 0) and c.flat().nnz == 10446320
    packed = tm.synth.load_packed("nsf")
    if packed is not None:
        np.testing.assert_array_equal(c.flat().N_cumsum, packed.N_cumsum)
        np.testing.assert_array_equal(c.flat().terms, packed.terms)
        np.testing.assert_array_equal(c.flat().counts, packed.counts)
    u = tm.readcorp(docfile=root + "/citeu/citeudocs.txt", vocabfile=root + "/citeu/citeuvocab.txt", userfile=root + "/citeu/citeuusers.txt",
                    counts=True, readers=True)
    assert u.size() == (16980, 8000, 5551) and u.flat().nnz == 1130920 and int(u.flat().R_cumsum[-1]) == 204986


def test_mapped_file_edges(tm, tmp_path):
    """The parser reads the file through a mapping: a file whose size is an exact multiple of the page size and that does not end in a
    newline must not be read past its end; an empty file is an empty corpus; a file of several MB is indexed by several threads and
    must give the same corpus as one thread."""
    import mmap
    p = str(tmp_path / "d.txt")
    page = mmap.PAGESIZE
    line = "1,2,3,4,5,6,7\n"
    body = line * (2 * page // len(line))
    pad = 2 * page - len(body)                     # fill up to exactly two pages with one last line without a newline
    if pad == 0:
        body, pad = body[:-len(line)], len(line)
    body += "1" * pad if pad <= 9 else "1," + "1" * (pad - 2)
    assert len(body) == 2 * page and not body.endswith("\n")
    open(p, "w").write(body)
    f = tm.readcorp(docfile=p).flat()
    assert f.M == body.count("\n") + 1 and f.nnz == body.count(",") + f.M
    open(p, "w").write("")
    assert tm.readcorp(docfile=p).flat().M == 0
    rng = np.random.default_rng(0)
    rows = ["%s\n%s\n" % (",".join(map(str, rng.integers(1, 50000, size=n))), ",".join(map(str, rng.integers(1, 9, size=n))))
            for n in rng.integers(1, 120, size=12000)]
    open(p, "w").write("".join(rows))
    assert os.path.getsize(p) > 4 << 20
    a, b = tm.readcorp(docfile=p, counts=True, nthreads=1).flat(), tm.readcorp(docfile=p, counts=True, nthreads=7).flat()
    for x, y in zip((a.N_cumsum, a.terms, a.counts), (b.N_cumsum, b.terms, b.counts)):
        np.testing.assert_array_equal(x, y)
    assert a.M == 12000


@pytest.mark.parametrize("body,kw,ok", [
    ("1,2\n1,1,1\n", dict(counts=True), False),          # more counts than terms
    ("1,2\n 1 ,\t+2\n", dict(counts=True), True),        # blanks and a sign: the general scan
    ("1,2\n1,-2\n", dict(counts=True), False),           # check_doc: counts must be positive
    ("2147483647\n", dict(), True),                      # the largest key the packed CSR holds
    ("2147483648\n", dict(), False),
    ("00000000001,2\n", dict(), True),                   # 11 digits: not the fast path, still 1
    ("1,,2\n", dict(), False),
    ("1,2,\n", dict(), False),
])
def test_field_parser_paths(tm, tmp_path, body, kw, ok):
    p = str(tmp_path / "d.txt")
    open(p, "w").write(body)
    if ok:
        f = tm.readcorp(docfile=p, **kw).flat()
        assert f.M == 1 and f.terms[0] + 1 == int(body.split("\n")[0].split(",")[0])
        if kw:
            np.testing.assert_array_equal(f.counts, [1, 2])
    else:
        with pytest.raises(tm.CorpusError, match=r"document 1 beginning on line 1 failed to load\."):
            tm.readcorp(docfile=p, **kw)


def test_key_files_that_are_not_the_unit_range(tm, tmp_path):
    """check_corp's order (Corpus.jl:111-122): a used key outside the key set is reported before the unit-range violation."""
    p, v = str(tmp_path / "d.txt"), str(tmp_path / "v.txt")
    open(p, "w").write("1,3\n")
    open(v, "w").write("1\ta\n3\tc\n7\tg\n")                      # contains every used key, but is not 1:3
    with pytest.raises(tm.CorpusError, match="must form unit range"):
        tm.readcorp(docfile=p, vocabfile=v)
    open(p, "w").write("1,2\n")
    with pytest.raises(tm.CorpusError, match="term keys not found"):
        tm.readcorp(docfile=p, vocabfile=v)
    open(v, "w").write("0x1\ta\n2.0\tb\n\n3\tc\n3\tc again\n")    # prefixes, an integral float, a blank line, a repeated key
    assert tm.readcorp(docfile=p, vocabfile=v).V == 3
    open(v, "w").write("1\ta\nx\tb\n")
    with pytest.raises(tm.CorpusError, match="vocab keys must be positive integers"):
        tm.readcorp(docfile=p, vocabfile=v)
