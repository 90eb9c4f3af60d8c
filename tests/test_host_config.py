"""Host-side model specification (plspm.config) -- same rules and errors as the reference
(reference tests/test_config.py).  CPU only."""
import numpy as np
import pandas as pd
import pytest

import plspm.config as c
from oracle import plspm_oracle as orc
from plspm.mode import Mode
from plspm.scale import Scale
from plspm.scheme import Scheme


def three_lv_path():
    lvs = ["AGRI", "IND", "POLINS"]
    return pd.DataFrame([[0, 0, 0], [0, 0, 0], [1, 1, 0]], index=lvs, columns=lvs)


def frame():
    rng = np.random.default_rng(0)
    cols = ["gini", "farm", "rent", "gnpr", "labo", "inst", "ecks", "death", "demo"]
    return pd.DataFrame(rng.standard_normal((40, len(cols))) * 2 + 5, columns=cols)


def test_rejects_bad_path_matrix():
    with pytest.raises(TypeError):
        c.Config("hello")
    with pytest.raises(ValueError):
        c.Config(pd.DataFrame([[0, 0, 0]]))
    with pytest.raises(ValueError):
        c.Config(pd.DataFrame([[1, 1], [1, 1]]))
    with pytest.raises(ValueError):
        c.Config(pd.DataFrame([[1, 0], [2, 1]]))
    with pytest.raises(ValueError):
        c.Config(pd.DataFrame([[1, 0], [1, 1]], index=["A", "B"], columns=["C", "D"]))


def test_rejects_unknown_lv_and_duplicate_columns():
    config = c.Config(three_lv_path())
    with pytest.raises(ValueError):
        config.add_lv("NOPE", Mode.A, c.MV("x"))
    config.add_lv("AGRI", Mode.A, c.MV("gini"), c.MV("farm"))
    with pytest.raises(ValueError):
        config.add_lv("IND", Mode.A, c.MV("gini"))
    with pytest.raises(ValueError):
        config.add_lv("IND", Mode.A, c.MV("POLINS"))


def test_mode_mvs_and_filter_order():
    config = c.Config(three_lv_path())
    config.add_lv("POLINS", Mode.B, c.MV("inst"), c.MV("ecks"))
    config.add_lv("AGRI", Mode.A, c.MV("gini"), c.MV("farm"), c.MV("rent"))
    config.add_lv("IND", Mode.A, c.MV("gnpr"))
    assert config.mode("AGRI") == Mode.A and config.mode("POLINS") == Mode.B
    assert config.mvs("AGRI") == ["gini", "farm", "rent"]
    # filtered columns keep add_lv insertion order (config.py:269); ODM rows follow the path order
    assert list(config.filter(frame())) == ["inst", "ecks", "gini", "farm", "rent", "gnpr"]
    odm = config.odm(config.path())
    assert list(odm.index) == ["gini", "farm", "rent", "gnpr", "inst", "ecks"]
    assert list(odm.columns) == ["AGRI", "IND", "POLINS"]
    assert odm.loc["gnpr", "IND"] == 1 and odm.loc["gnpr", "AGRI"] == 0


def test_filter_errors():
    config = c.Config(three_lv_path())
    config.add_lv("AGRI", Mode.A, c.MV("gini"))
    config.add_lv("IND", Mode.A, c.MV("gnpr"))
    with pytest.raises(ValueError):  # POLINS not configured
        config.filter(frame())
    config.add_lv("POLINS", Mode.A, c.MV("absent"))
    with pytest.raises(ValueError):  # column not in data
        config.filter(frame())
    config2 = c.Config(three_lv_path())
    config2.add_lv("AGRI", Mode.A, c.MV("gini"))
    config2.add_lv("IND", Mode.A, c.MV("gnpr"))
    config2.add_lv("POLINS", Mode.A, c.MV("inst"))
    bad = frame()
    bad["gini"] = bad["gini"].astype(str)
    with pytest.raises(ValueError):
        config2.filter(bad)


def test_filter_drops_rows_with_a_fully_missing_block():
    config = c.Config(three_lv_path())
    config.add_lv("AGRI", Mode.A, c.MV("gini"), c.MV("farm"))
    config.add_lv("IND", Mode.A, c.MV("gnpr"))
    config.add_lv("POLINS", Mode.A, c.MV("inst"))
    df = frame()
    df.loc[3, ["gini", "farm"]] = np.nan
    df.loc[7, "gini"] = np.nan
    out = config.filter(df)
    assert 3 not in out.index and 7 in out.index and config.missing()


def test_scale_bookkeeping_of_nonmetric_configs():
    config = c.Config(three_lv_path())
    config.add_lv("AGRI", Mode.A, c.MV("gini", Scale.NUM), c.MV("farm"))
    config.add_lv("IND", Mode.A, c.MV("gnpr"))
    config.add_lv("POLINS", Mode.A, c.MV("inst"))
    assert not config.metric()
    with pytest.raises(TypeError):
        config.treat(config.filter(frame()))
    allraw = c.Config(three_lv_path(), default_scale=Scale.RAW)
    allraw.add_lv("AGRI", Mode.A, c.MV("gini"))
    allraw.add_lv("IND", Mode.A, c.MV("gnpr", Scale.NUM))
    allraw.add_lv("POLINS", Mode.A, c.MV("inst"))
    df = allraw.filter(frame())
    treated = allraw.treat(df)  # config.py:314: unit POPULATION variance per column
    np.testing.assert_allclose(treated.to_numpy().std(axis=0, ddof=0), 1.0, rtol=1e-12)
    np.testing.assert_allclose(treated.to_numpy().mean(axis=0), 0.0, atol=1e-12)
    assert allraw.scaled() and allraw.scale("gini") == Scale.NUM and allraw.numeric()
    ordinal = c.Config(three_lv_path(), default_scale=Scale.ORD)
    ordinal.add_lv("AGRI", Mode.A, c.MV("gini"))
    ordinal.add_lv("IND", Mode.A, c.MV("gnpr", Scale.NUM))
    ordinal.add_lv("POLINS", Mode.A, c.MV("inst"))
    assert not ordinal.numeric()
    with pytest.raises(NotImplementedError):
        ordinal.treat(ordinal.filter(frame()))


def test_structure_builds_lower_triangular_path_in_reference_order():
    lvs = ["MANDRILL", "BONOBO", "APE", "GOAT", "CATFISH"]
    expected = pd.DataFrame([[0, 0, 0, 0, 0], [0, 0, 0, 0, 0], [1, 1, 0, 0, 0], [0, 0, 1, 0, 0], [0, 0, 1, 0, 0]],
                            index=lvs, columns=lvs)
    s = c.Structure()
    s.add_path(source=["BONOBO", "MANDRILL"], target=["APE"])
    s.add_path(source=["APE"], target=["CATFISH", "GOAT"])
    pd.testing.assert_frame_equal(expected, s.path())
    pd.testing.assert_frame_equal(expected, s.path())  # not single-use (quirk Q5 fixed)
    with pytest.raises(ValueError):
        s.add_path(["A", "B"], ["C", "D"])
    cyc = c.Structure()
    cyc.add_path(["A"], ["B"])
    cyc.add_path(["B"], ["A"])
    with pytest.raises(ValueError):
        cyc.path()
    again = c.Structure(expected).path()  # same edges; LV order follows insertion
    assert all(again.loc[r, col] == expected.loc[r, col] for r in lvs for col in lvs)


def test_satisfaction_structure_order(sat):
    s = c.Structure()
    s.add_path(["IMAG"], ["EXPE", "SAT", "LOY"])
    s.add_path(["EXPE"], ["QUAL", "VAL", "SAT"])
    s.add_path(["QUAL"], ["VAL", "SAT"])
    s.add_path(["VAL"], ["SAT"])
    s.add_path(["SAT"], ["LOY"])
    assert list(s.path()) == [str(v) for v in sat["lvs"]]
    np.testing.assert_array_equal(s.path().values, sat["path"])


@pytest.mark.parametrize("scaled", (False, True))
def test_metric_treat_matches_oracle(scaled):
    config = c.Config(three_lv_path(), scaled=scaled)
    config.add_lv("AGRI", Mode.A, c.MV("gini"), c.MV("farm"))
    config.add_lv("IND", Mode.A, c.MV("gnpr"))
    config.add_lv("POLINS", Mode.A, c.MV("inst"))
    df = frame()
    df.loc[5, "farm"] = np.nan
    f = config.filter(df)
    np.testing.assert_allclose(config.treat(f).values, orc.treat_metric(f.values, scaled), rtol=1e-12, atol=1e-14)


def test_clone_is_independent():
    config = c.Config(three_lv_path())
    config.add_lv("AGRI", Mode.A, c.MV("gini"))
    other = config.clone()
    other.add_lv("IND", Mode.B, c.MV("gnpr"))
    with pytest.raises(KeyError):
        config.mode("IND")
    assert other.mode("IND") == Mode.B and other.mvs("AGRI") == ["gini"]


def test_enums():
    assert {s.name for s in Scheme} == {"CENTROID", "PATH", "FACTORIAL"}
    assert Mode.A != Mode.B and Mode.A.value.engine_id == 0 and Mode.B.value.engine_id == 1
    assert Scheme.CENTROID.value.engine_id == 0 and Scheme.FACTORIAL.value.engine_id == 1
    assert Scheme.PATH.value.engine_id == 2


def test_hoc_first_stage_path_matches_reference(syn):
    from plspm.estimator import Estimator
    if "hoc/first_stage_path" not in syn.files:
        pytest.skip("fixture without the HOC case")
    st = c.Structure()
    st.add_path(["Expectation", "Quality"], ["Satisfaction"])
    st.add_path(["Satisfaction"], ["Complaints", "Loyalty"])
    cfg = c.Config(st.path())
    cfg.add_higher_order("Satisfaction", Mode.A, ["Image", "Value"])
    fs = Estimator(cfg).hoc_path_first_stage(cfg)
    assert list(fs.index) == [str(v) for v in syn["hoc/first_stage_lvs"]]
    np.testing.assert_array_equal(fs.values, syn["hoc/first_stage_path"])
    assert "Satisfaction" not in fs.index and {"Image", "Value"} <= set(fs.index)
    with pytest.raises(NotImplementedError):
        Estimator(cfg).estimate(type("Calc", (), {"config": lambda self: cfg})(), None)
