"""The JVM-side veneer (jni/marlin_b200_jni.c + scala/...) cannot run here (no JDK, no scalac), so it is checked the
way a build without a JVM can: the C file is parsed and type-checked by gcc against the real C ABI header (with a
declaration-only stand-in for <jni.h>), every `@native` in Native.scala has its JNI twin with the same arity, and every
C-ABI entry the veneer calls is declared in include/marlin_b200.h."""
import re
import shutil
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
JNI = ROOT / "jni" / "marlin_b200_jni.c"
NATIVE = ROOT / "scala" / "edu" / "nju" / "pasalab" / "marlin" / "matrix" / "Native.scala"


def test_jni_translation_unit_type_checks_against_the_c_abi():
    gcc = shutil.which("gcc")
    assert gcc
    out = subprocess.run([gcc, "-std=c11", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter", "-fsyntax-only", f"-I{ROOT / 'include'}",
                          f"-I{ROOT / 'jni'}", str(JNI)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert out.returncode == 0, out.stdout


def _scala_natives():
    text = NATIVE.read_text()
    body = text[text.index("object Native"):text.index("object Ctx")]
    body = re.sub(r"//[^\n]*", "", body)
    out = {}
    for mm in re.finditer(r"@native\s+def\s+(\w+)\s*\(([^)]*)\)\s*:\s*([\w\[\]]+)", body, re.S):
        params = [p for p in mm.group(2).split(",") if p.strip()]
        out[mm.group(1)] = len(params)
    return out


def _c_natives():
    text = JNI.read_text()
    out = {}
    for mm in re.finditer(r"NATIVE\((\w+), (\w+)\)\(JNIEnv\* e, jobject o([^)]*)\)", text):
        params = [p for p in mm.group(3).split(",") if p.strip()]
        out[mm.group(2)] = len(params)
    return out


def test_every_scala_native_has_its_jni_twin():
    scala, c = _scala_natives(), _c_natives()
    assert len(scala) >= 40
    assert set(scala) == set(c), set(scala) ^ set(c)
    for name, arity in scala.items():
        assert c[name] == arity, (name, arity, c[name])


def test_veneer_only_calls_declared_abi_entries():
    header = (ROOT / "include" / "marlin_b200.h").read_text()
    declared = set(re.findall(r"\b(mb_\w+)\s*\(", header))
    called = set(re.findall(r"\b(mb_\w+)\s*\(", JNI.read_text()))
    assert called <= declared, called - declared
    for name in ("mb_block_gemm", "mb_block_add", "mb_block_transpose", "mb_matmul_blocked_dist", "mb_comm_init", "mb_choose_split"):
        assert name in called


def test_scala_submatrix_keeps_the_reference_method_surface():
    """matrix/SubMatrix.scala:27-139: rows, cols, isSparse, add/subtract (block and scalar), divide, multiply overloads."""
    text = (ROOT / "scala" / "edu" / "nju" / "pasalab" / "marlin" / "matrix" / "SubMatrix.scala").read_text()
    for sig in ("def this(denseMatrix: BDM[Double])", "def isSparse", "def add(other: SubMatrix): SubMatrix", "def add(b: Double): SubMatrix",
                "def subtract(other: SubMatrix): SubMatrix", "def subtract(b: Double): SubMatrix", "def divide(b: Double): SubMatrix",
                "def multiply(other: SubMatrix): SubMatrix", "def multiply(other: BDM[Double]): SubMatrix", "def multiply(b: Double): SubMatrix"):
        assert sig in text, sig
    assert "val rows: Int" in text and "val cols: Int" in text


def test_scala_sources_only_use_declared_natives():
    """Every `Native.x` used by the Scala host sources (SubMatrix, BlockMatrixMultiply, DenseVecMatrixNative, NativeSplit,
    NativeRandom) is an `@native def` (or a constant) of Native.scala — the closest thing to a link check without scalac."""
    natives = set(_scala_natives())
    text = NATIVE.read_text()
    consts = set(re.findall(r"\bval\s+(\w+)\s*=", text[text.index("object Native"):text.index("object Ctx")]))
    used = {}
    for f in (ROOT / "scala").rglob("*.scala"):
        for name in re.findall(r"\bNative\.(\w+)", f.read_text()):
            used.setdefault(name, f.name)
    missing = {n: f for n, f in used.items() if n not in natives | consts}
    assert not missing, missing
    for name in ("matmulRowshardedHost", "lu", "cholesky", "inverse", "trsm", "fillUniform", "partitionSeeds", "matmulBlockedDist"):
        assert name in used, name
