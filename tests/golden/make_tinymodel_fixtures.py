"""tests/golden/tinymodel.json.gz + models/tinymodel*.bin.gz: the reference's own known-answer test for the NN path
(cpp/tests/tinymodel.cpp): its two embedded tiny nets, its three test positions, the outputs it expects and the tolerances it
allows.  The numbers are read from the reference's test source, the nets are decoded from its base64 constants, the input rows
come from the reference's fillRowV7 (`kgref_driver tinyfeatures`)."""
import base64, gzip, json, os, re, subprocess, tempfile
HERE = os.path.dirname(os.path.abspath(__file__))
DRIVER = os.path.join(HERE, "..", "..", "oracle", "_ref", "kgref_driver")
REF = "/root/reference/cpp/tests"
src = open(os.path.join(REF, "tinymodel.cpp")).read()
data = open(os.path.join(REF, "tinymodeldata.cpp")).read() if os.path.exists(os.path.join(REF, "tinymodeldata.cpp")) else src


def const_string(name):
    m = re.search(r"const char\*\s+TinyModelTest::" + name + r"\s*=\s*R\"%%\((.*?)\)%%\"", data, re.S)
    return "".join(m.group(1).split())


tiny = base64.b64decode("".join(const_string(f"tinyModelBase64Part{i}") for i in range(7)))
mish = base64.b64decode(const_string("tinyMishModelBase64"))
os.makedirs(os.path.join(HERE, "models"), exist_ok=True)
open(os.path.join(HERE, "models", "tinymodel.bin.gz"), "wb").write(tiny)
open(os.path.join(HERE, "models", "tinymishmodel.bin.gz"), "wb").write(mish)

blocks = []
for m in re.finditer(r"setDefaultSymmetry\((\d)\);\s*Board board = Board::parseBoard\((\d+),(\d+),R\"%%\((.*?)\)%%\"\);(.*?)runOneTest\(\);", src, re.S):
    sym, X, Y, diagram, body = int(m.group(1)), int(m.group(2)), int(m.group(3)), m.group(4), m.group(5)
    scalars = {k: (float(v), float(t)) for k, v, t in re.findall(r"EQ\(nnOutput\.(\w+),\s*([-\d.]+),\s*([\d.]+)\);", body)}
    arrays = {k: [float(x) for x in re.findall(r"-?\d+(?:\.\d+)?", v)] for k, v in re.findall(r"double (expected\w+)\[[^\]]*\] = \{(.*?)\};", body, re.S)}
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        f.write(diagram)
    feat = json.loads(subprocess.run([DRIVER, "tinyfeatures", f.name, str(X), str(Y)], capture_output=True, text=True, check=True).stdout)
    os.unlink(f.name)
    loops = re.findall(r"EQ\((\w+)\*10000, (expected\w+)\[(?:pos|idx)\], (.*?)\);", body)
    blocks.append(dict(model="tinymodel" if len(blocks) == 0 else "tinymishmodel", symmetry=sym, X=X, Y=Y, scalars=scalars, arrays=arrays,
                       tolerance_exprs={name: expr for _, name, expr in loops}, **feat))
    print(blocks[-1]["model"], sym, X, Y, sorted(scalars), {k: len(v) for k, v in arrays.items()}, blocks[-1]["tolerance_exprs"])
with gzip.GzipFile(os.path.join(HERE, "tinymodel.json.gz"), "wb", mtime=0) as f:
    f.write(json.dumps(blocks).encode())
print(len(tiny), len(mish), "model bytes")
