"""Product-side Gmsh reader (lehrfempp_b200/csrc/gmsh.cpp, host code behind the C ABI) against the oracle's restatement of
lf::io::GmshReader on the reference's own input files -- no GPU needed: the reader is host logic."""
import os

import numpy as np
import pytest

import lehrfempp_b200 as lf
from oracle.lfo_gmsh import GmshError
from oracle.lfo_gmsh import GmshReader as OracleReader

MSH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "msh")
FILES = sorted(n for n in os.listdir(MSH) if n.endswith(".msh"))


def test_fixture_set_is_complete():
    assert len(FILES) == 19


@pytest.mark.parametrize("name", FILES)
def test_reader_matches_oracle(name):
    path = os.path.join(MSH, name)
    o = OracleReader(path)
    g = lf.GmshReader(path)
    oxy, oen, ocn, oorder = o.arrays()
    gxy, gen, gcn = g.arrays()
    # numbering handed to the MeshFactory: bit-exact
    assert np.array_equal(gxy, oxy) and np.array_equal(gen, oen) and np.array_equal(gcn, ocn)
    assert g.geometry_order == oorder
    # physical entity numbers of every node, explicit edge and cell (+ a few indices beyond: no numbers)
    for codim, n in ((2, g.n_nodes), (1, g.n_explicit_edges), (0, g.n_cells)):
        for i in range(n + 3):
            assert g.physical_entity_nr(codim, i) == o.physical_entity_nr(codim, i)
        assert g.physical_entities(codim) == o.physical_entities(codim)
        numbers = {nr for i in range(n) for nr in o.physical_entity_nr(codim, i)} | {99}
        for nr in numbers:
            want = np.array([o.is_physical_entity(codim, i, nr) for i in range(n + 2)], dtype=np.uint8)
            assert np.array_equal(g.physical_flags(codim, nr, n + 2), want)
    # name <-> number tables, including the ambiguous and the missing cases
    for nr, name_, cd in o.names:
        assert g.name2nr(name_, cd) == o.name2nr(name_, cd) and g.nr2name(nr, cd) == o.nr2name(nr, cd)
        for fn_o, fn_g, arg in ((o.name2nr, g.name2nr, name_), (o.nr2name, g.nr2name, nr)):
            try:
                want = fn_o(arg)
            except GmshError:
                with pytest.raises(lf.LfgpuError):
                    fn_g(arg)
            else:
                assert fn_g(arg) == want
    with pytest.raises(lf.LfgpuError):
        g.name2nr("gugus")
    with pytest.raises(lf.LfgpuError):
        g.nr2name(100)


def test_read_from_memory_equals_read_from_file():
    path = os.path.join(MSH, "two_element_hybrid_2d_v4_binary.msh")
    a = lf.GmshReader(path)
    with open(path, "rb") as fh:
        b = lf.GmshReader(fh.read())
    for u, v in zip(a.arrays(), b.arrays()):
        assert np.array_equal(u, v)


def test_error_behaviour():
    with pytest.raises(lf.LfgpuError) as e:
        lf.GmshReader(os.path.join(MSH, "does_not_exist.msh"))
    assert "Could not open file" in str(e.value)  # gmsh_reader.cc:633-637
    with pytest.raises(lf.LfgpuError) as e:
        lf.GmshReader(b"$MeshFormat\n3.0 0 8\n$EndMeshFormat\n")
    assert "not yet supported" in str(e.value)  # gmsh_reader.cc:693-694
    with pytest.raises(lf.LfgpuError) as e:
        lf.GmshReader(b"$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n3\n1 0 0 0\n2 1 0 0\n3 0 1 1\n$EndNodes\n"
                      b"$Elements\n1\n1 2 2 1 1 1 2 3\n$EndElements\n")
    assert "z-coordinate" in str(e.value)  # gmsh_reader.cc:197-199
    with pytest.raises(lf.LfgpuError) as e:
        lf.GmshReader(b"$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n2\n1 0 0 0\n2 1 0 0\n$EndNodes\n"
                      b"$Elements\n1\n1 1 2 1 1 1 2\n$EndElements\n")
    assert "no elements with dimension 2" in str(e.value)  # gmsh_reader.cc:172-173
    with pytest.raises(lf.LfgpuError):
        lf.GmshReader(b"$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n4\n1 0 0 0\n2 1 0 0\n3 0 1 0\n4 0 0 1\n$EndNodes\n"
                      b"$Elements\n1\n1 4 2 1 1 1 2 3 4\n$EndElements\n")  # a tetrahedron in a 2D mesh
    with pytest.raises(lf.LfgpuError):
        lf.GmshReader(b"$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n3\n1 0 0 0\n")  # truncated


def test_non_consecutive_repetition_is_not_merged():
    """Only CONSECUTIVE repetitions of an element are merged (gmsh_reader.cc:224-229)."""
    txt = (b"$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n4\n1 0 0 0\n2 1 0 0\n3 1 1 0\n4 0 1 0\n$EndNodes\n$Elements\n5\n"
           b"1 1 2 7 1 1 2\n2 1 2 8 1 1 2\n3 1 2 9 1 2 3\n4 2 2 5 1 1 2 3\n5 2 2 6 1 1 3 4\n$EndElements\n")
    g, o = lf.GmshReader(txt), OracleReader(txt)
    assert g.n_explicit_edges == 2 and g.n_cells == 2
    assert g.physical_entity_nr(1, 0) == [7, 8] == o.physical_entity_nr(1, 0)
    assert g.physical_entity_nr(1, 1) == [9] and g.physical_entity_nr(0, 1) == [6]


def test_shim_reader_passes_the_reference_reader_tests():
    """tests/cpp/gmsh_shim_test.cc: checkTwoElementMesh / checkPieceOfCake of gmsh_reader_tests.cc against lfgpu::GmshReader
    (the C++ shim class with the reference's member names)."""
    import subprocess
    cpp = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp")
    subprocess.check_call(["make", "-C", cpp, "-s", "gmsh_shim_test"])
    out = subprocess.run([os.path.join(cpp, "gmsh_shim_test"), MSH], capture_output=True, text=True)
    assert out.returncode == 0 and "GMSH_SHIM_TEST_OK" in out.stdout, out.stdout + out.stderr


def test_corrupt_counts_are_rejected_not_allocated():
    """Counts larger than the file can hold must end in an error, not in an allocation of that size (the parser was
    fuzzed with 40000 mutated copies of the fixtures under ASan/UBSan)."""
    for txt in (b"$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n99999999999999999\n1 0 0 0\n$EndNodes\n",
                b"$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n1\n1 0 0 0\n$EndNodes\n$Elements\n1\n1 2 99999999999 1 1 1 1 1\n$EndElements\n",
                b"$MeshFormat\n4.1 0 8\n$EndMeshFormat\n$Nodes\n1 99999999999999 1 1\n2 1 0 1\n1\n0 0 0\n$EndNodes\n",
                b"$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n-5\n$EndNodes\n"):
        with pytest.raises(lf.LfgpuError):
            lf.GmshReader(txt)
    with pytest.raises(lf.LfgpuError) as e:
        lf.GmshReader(b"$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n1\n1 0 \xf5\xff 0\n$EndNodes\n")
    assert "expected a number" in str(e.value)  # file bytes quoted in the message are made printable
