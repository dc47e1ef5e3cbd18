"""Turns the reference's own ``spatial_image_analysis.py`` (Python 2) into a module Python 3 can import, WITHOUT copying
it into the repository: the source is read where it lies under ``/root/reference``, rewritten by the mechanical rules
below and written to ``oracle/_ref/vplants_ref/spatial_image_analysis.py`` (git-ignored; built by ``build()``).

TEST INFRASTRUCTURE.  ``tests/test_reference_pin.py`` imports the result next to ``oracle/sia_loops.py`` and asserts that
the restatement equals the reference itself on the docstring image, Voronoi domes and a label-0 volume -- that is what
pins the oracle beyond the reference's seven docstring vectors.  ``/root/reference`` does not exist on the GPU box;
there the test uses the module built here (``oracle/_ref`` travels with the snapshot) or skips.

Rewrite rules (syntax only, plus the three numpy >= 2 incompatibilities SURVEY.md section 8c lists):
  print statements -> print();  xrange -> range;  d.has_key(k) -> (k in d);  iteritems / iterkeys / itervalues;
  ``x.itervalues().next()`` -> ``next(iter(x.values()))``;  ``cPickle`` -> ``pickle``;  ``np.bool`` -> ``np.bool_``;
  SIA:1231  ``index=np.int16(labels)``   numpy 2 raises OverflowError above 32767 (numpy 1 wrapped, giving volume 0)
            -> ``index=np.asarray(labels)``  (the wide index the oracle documents as its one deviation);
  SIA:1033, 1441  boolean ``a - b`` (a TypeError today) -> ``a ^ b`` where b is a subset of a, as it is at both sites;
  SIA:861   ``if x != []`` on an ndarray (element-wise today) -> ``if len(x) != 0``;
  SIA:37, 42  ``dilation`` / ``dilation_by`` return a LIST of slices, an index numpy >= 1.23 rejects (the reference's bare
            ``except:`` would then silently use the whole image) -> a tuple, what the numpy of its time made of it;
  SIA:514   ``nd.find_objects(self.image==0)``: today's scipy rejects a boolean input -> ``.astype(np.uint8)``;
  SIA:1000  ``map(int, l)`` -> ``list(map(int, l))``.
``openalea.image.serial.basics`` (absent here) is replaced by ``oracle/ref_stubs.py`` through ``sys.modules``.
"""
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src/vplants/tissue_analysis/spatial_image_analysis.py"
OUT_DIR = os.path.join(HERE, "_ref", "vplants_ref")


def _split_statement(rest):
    """rest of a line after ``print`` -> (arguments, tail) where tail starts at an unquoted ';' or '#'."""
    depth = 0
    quote = None
    i = 0
    while i < len(rest):
        c = rest[i]
        if quote:
            if c == "\\":
                i += 2
                continue
            if rest.startswith(quote, i):
                i += len(quote)
                quote = None
                continue
        else:
            if rest.startswith('"""', i) or rest.startswith("'''", i):
                quote = rest[i:i + 3]
                i += 3
                continue
            if c in "\"'":
                quote = c
            elif c in "([{":
                depth += 1
            elif c in ")]}":
                depth -= 1
            elif depth == 0 and c in ";#":
                return rest[:i], rest[i:]
        i += 1
    return rest, ""


_PRINT = re.compile(r"^(?P<head>\s*(?:(?:if|elif|while|for)\b[^:]*:\s*|else:\s*|try:\s*|except[^:]*:\s*|finally:\s*)?)print(?P<rest>(?:\s.*)?)$")


def _convert_print(line):
    body = line.rstrip("\n")
    m = _PRINT.match(body)
    if not m or body.lstrip().startswith("#"):
        return line
    head, rest = m.group("head"), m.group("rest") or ""
    args, tail = _split_statement(rest)
    args = args.strip()
    if args.startswith("("):          # already a call (none in the reference today)
        return line
    end = ""
    if args.endswith(","):
        args = args[:-1].rstrip()
        end = ", end=' '" if args else "end=' '"
    new = "%sprint(%s%s)%s" % (head, args, end, (" " if tail and not tail.startswith(";") else "") + tail if tail else "")
    return new + "\n"


def convert(src):
    out = []
    for line in src.splitlines(True):
        if "print" in line:
            # ``if verbose: print "a",; percent += 5`` -> the print is one statement of several
            line = _convert_print(line)
        out.append(line)
    s = "".join(out)
    s = s.replace("cPickle as pickle", "pickle as pickle")
    s = re.sub(r"\bxrange\(", "range(", s)
    s = re.sub(r"(\w+)\.itervalues\(\)\.next\(\)", r"next(iter(\1.values()))", s)
    s = s.replace(".iteritems()", ".items()").replace(".iterkeys()", ".keys()").replace(".itervalues()", ".values()")
    s = re.sub(r"([\w\.]+)\.has_key\(", r"\1.__contains__(", s)
    s = re.sub(r"\bnp\.bool\b(?!_)", "np.bool_", s)
    # numpy >= 2
    s = s.replace("index=np.int16(labels)", "index=np.asarray(labels)")
    s = s.replace("if x != []", "if len(x) != 0")
    s = s.replace("layer = dil_1 - mask_img_1", "layer = dil_1 ^ mask_img_1")
    s = s.replace("mask_bbox_im - eroded_mask_bbox_im", "mask_bbox_im ^ eroded_mask_bbox_im")
    # a LIST of slices as an index (SIA:37, 42 feed SIA:600, 786, 831, 933, ...) was read as a tuple by the numpy of the
    # reference's time; numpy >= 1.23 raises, which the reference's bare ``except:`` turns into "use the whole image"
    for old in ("return [ slice(max(0,s.start-1), s.stop+1) for s in slices ]",
                "return [ slice(max(0,s.start-amount), s.stop+amount) for s in slices ]"):
        assert old in s, old
        s = s.replace(old, "return tuple(" + old[len("return "):] + ")")
    # scipy today rejects a boolean input of find_objects (SIA:514)
    s = s.replace("nd.find_objects(self.image==0)[0]", "nd.find_objects((self.image==0).astype(np.uint8))[0]")
    # py2 map() returns a list (SIA:1000 sorts and indexes it)
    s = s.replace("integers = lambda l : map(int, l)", "integers = lambda l : list(map(int, l))")
    return s


def build(ref_src=REF_SRC, out_dir=OUT_DIR, quiet=True):
    """Writes the importable module; returns its path, or None when the reference source is not there."""
    if not os.path.exists(ref_src):
        return None
    with open(ref_src, encoding="utf-8") as f:
        src = f.read()
    py3 = convert(src)
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "__init__.py"), "w") as f:
        f.write("# generated by oracle/make_ref.py from %s -- not tracked\n" % ref_src)
    path = os.path.join(out_dir, "spatial_image_analysis.py")
    with open(path, "w", encoding="utf-8") as f:
        f.write("# GENERATED by oracle/make_ref.py from %s (mechanical py2 -> py3 rewrite); do not edit, do not commit\n" % ref_src)
        f.write(py3)
    compile(py3, path, "exec")          # a syntax error here is a missing rewrite rule
    if not quiet:
        print("wrote", path)
    return path


def load():
    """Imports the generated module (building it first when the reference source is present).  None if unavailable."""
    import importlib
    path = os.path.join(OUT_DIR, "spatial_image_analysis.py")
    if os.path.exists(REF_SRC):
        build()
    if not os.path.exists(path):
        return None
    from . import ref_stubs
    ref_stubs.install()
    parent = os.path.dirname(OUT_DIR)
    if parent not in sys.path:
        sys.path.insert(0, parent)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):          # "PlantGL is not installed ..." (SIA:32)
        return importlib.import_module("vplants_ref.spatial_image_analysis")


if __name__ == "__main__":
    p = build(quiet=False)
    sys.exit(0 if p else 1)
