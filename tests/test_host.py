"""CPU: host-side mirror of the reference structs (src/all_structs.jl:62-72, :564-568; src/genomes/genomes.jl:150-175;
src/grm/grm.jl:27-29, :55-61, :94-96, :131-141) and argument validation that happens before the GPU is touched."""
import numpy as np
import pytest

import grm_b200
from grm_b200 import GRM, ArgumentError, Genomes, checkdims, clone, grmploidyaware, grmsimple, inflatediagonals, synth


def test_grm_struct_clone_hash_eq_checkdims():
    # grm.jl:18-25, :48-53, :80-92, :119-129 — hand-built 2 x 2 GRMs
    g = GRM(["entry_1", "entry_2"], ["l1", "l2", "l3"], np.array([[1.0, 0.5], [0.5, 1.0]]))
    c = clone(g)
    assert c == g and hash(c) == hash(g) and c is not g
    c.genomic_relationship_matrix[0, 0] = 2.0
    assert c != g and g.genomic_relationship_matrix[0, 0] == 1.0
    assert checkdims(g)
    g.entries = ["dummy_entry"]
    assert not checkdims(g)


def test_genomes_checkdims_rules():
    A = synth.dosage(4, 6, 2, seed=1)
    g = Genomes.from_matrix(A)
    assert checkdims(g)
    assert g.entries == ["entry_1", "entry_2", "entry_3", "entry_4"]
    bad = Genomes.from_matrix(A)
    bad.entries[1] = bad.entries[0]  # duplicated entry names
    assert not checkdims(bad)
    bad = Genomes.from_matrix(A)
    bad.loci_alleles = bad.loci_alleles[:-1]
    assert not checkdims(bad)
    bad = Genomes.from_matrix(A)
    bad.mask = np.ones((4, 5), dtype=bool)
    assert not checkdims(bad)


@pytest.mark.parametrize("fn", [grmsimple, grmploidyaware])
def test_corrupted_genomes_raise_argument_error_before_any_gpu_work(fn):
    g = Genomes.from_matrix(synth.dosage(4, 6, 2, seed=1))
    g.populations = g.populations[:-1]
    with pytest.raises(ArgumentError, match="The Genomes struct is corrupted"):  # calc.jl:118-120, :191-193
        fn(g)


def test_inflatediagonals_rejects_non_square_as_not_symmetric():
    with pytest.raises(ArgumentError, match="not symmetric"):  # calc.jl:43 runs before the squareness check
        inflatediagonals(np.zeros((2, 3)))


def test_synthetic_generators_have_the_simulator_shape():
    A = synth.continuous(50, 400, seed=42)
    assert A.flags.f_contiguous and A.min() >= 0.0 and A.max() <= 1.0
    assert (A == 0.0).any() and (A == 1.0).any()  # clipping atoms, simulate_genomes.jl:515-516
    D = synth.dosage(50, 400, 4, seed=42)
    assert set(np.unique(D)) <= {0.0, 0.25, 0.5, 0.75, 1.0}
    M = synth.add_missing(D, 0.1, seed=1)
    assert np.isnan(M).sum() == round(0.1 * 50 * 400)  # simulate_genomes.jl:738


def test_distances_validates_like_the_reference_before_any_gpu_work():
    # genomes.jl:463-500 — corrupted struct, metric list, index range; all raised before the backend is touched
    g = Genomes.from_matrix(synth.continuous(6, 120, seed=3))
    with pytest.raises(ArgumentError, match="Unrecognised metric/s"):
        grm_b200.distances(g, ["euclidean", "covariance"], [1, 2])   # "covariance" is not recognised (:470)
    with pytest.raises(ArgumentError, match="at least 1 distance metric"):
        grm_b200.distances(g, [], [1, 2])
    with pytest.raises(ArgumentError, match="from 1 to `n_loci_alleles`"):
        grm_b200.distances(g, ["mad"], [0, 5])
    with pytest.raises(ArgumentError, match="from 1 to `n_loci_alleles`"):
        grm_b200.distances(g, ["mad"], [1, 121])
    small = Genomes.from_matrix(synth.continuous(6, 50, seed=3))
    with pytest.raises(ArgumentError, match="sample of 100"):
        grm_b200.distances(small, ["mad"])                            # sample(1:50, 100, replace = false) throws too
    g.populations = g.populations[:-1]
    with pytest.raises(ArgumentError, match="corrupted"):
        grm_b200.distances(g, ["mad"], [1, 2])


def test_simulatecovariancekinship_checks_the_entry_count_first():
    g = Genomes.from_matrix(synth.continuous(6, 120, seed=3))
    with pytest.raises(ArgumentError, match="p must match the number of entries"):  # simulate_effects.jl:345-347
        grm_b200.simulatecovariancekinship(7, g)


def test_product_package_never_imports_the_oracle():
    # the oracle is test infrastructure: nothing under the package may import, load or execute it
    import pathlib
    import re

    pkg = pathlib.Path(grm_b200.__file__).parent
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|libgrm_oracle|grm_oracle", re.M)
    for path in list(pkg.glob("*.py")) + list((pkg / "csrc").glob("*")):
        if path.is_file() and path.suffix in (".py", ".cu", ".cuh", ".cpp", ".h", "") and path.name != "Makefile":
            assert not pat.search(path.read_text(errors="ignore")), path
