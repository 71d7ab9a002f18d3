"""Rows written by ``extract.write_lmdb`` are read back by the reference's OWN reader:
``LMDBDataset.__getitem__`` (utils/datasets/lmdb_dataset.py:79-89), imported unmodified from
/root/reference or the staged baseline/_ref copy.  The ``lmdb`` package is not in this image;
when it is importable the test uses it, otherwise a dict-backed stand-in with the handful of
calls the reference makes (open / open_db / begin / put / get / cursor / stat)."""
import importlib
import pickle
import sys
import types

import numpy as np
import pytest
import torch

from interactive_spectrogram_inpainting_b200 import extract
from oracle import ref_loader

REF_MODULE = "interactive_spectrogram_inpainting.utils.datasets.lmdb_dataset"


class _FakeLmdb(types.ModuleType):
    """The subset of py-lmdb the reference touches, over per-path dicts."""

    def __init__(self):
        super().__init__("lmdb")
        self.stores = {}
        outer = self

        class Cursor:
            def __init__(self, table): self.table = table
            def first(self): return bool(self.table)
            def iternext(self, keys=True, values=True):
                for k in sorted(self.table):
                    yield k if not values else (k, self.table[k])

        class Txn:
            def __init__(self, env, db, write): self.env, self.db, self.write = env, db, write
            def __enter__(self): self.env.transactions += 1; return self
            def __exit__(self, *exc): return False
            def _table(self, db=None): return self.env.tables[db if db is not None else self.db]
            def put(self, key, value): assert self.write; self._table()[key] = value; return True
            def get(self, key): return self._table().get(key)
            def cursor(self): return Cursor(self._table())
            def stat(self, db=None): return {"entries": len(self._table(db))}

        class Env:
            def __init__(self): self.tables, self.transactions = {None: {}}, 0
            def open_db(self, name, dupsort=False): self.tables.setdefault(name, {}); return name
            def begin(self, db=None, write=False): return Txn(self, db, write)
            def __bool__(self): return True

        def open_(path, **kwargs):
            return outer.stores.setdefault(str(path), Env())
        self.open = open_
        self.Environment = Env


@pytest.fixture
def reference_lmdb_dataset():
    if not ref_loader.available():
        pytest.skip("neither /root/reference nor baseline/_ref is present")
    try:
        lmdb = importlib.import_module("lmdb")
        lent_lmdb = False
    except ImportError:
        lmdb = _FakeLmdb()
        sys.modules["lmdb"] = lmdb
        lent_lmdb = True
    before = set(sys.modules)
    root = str(ref_loader.REFERENCE_ROOT)
    added_path = root not in sys.path
    if added_path:
        sys.path.append(root)
    try:
        yield importlib.import_module(REF_MODULE), lmdb
    finally:
        for name in set(sys.modules) - before:
            if name.startswith("interactive_spectrogram_inpainting.utils"):
                del sys.modules[name]
        if lent_lmdb:
            del sys.modules["lmdb"]
        if added_path:
            sys.path.remove(root)


def test_reference_lmdb_dataset_reads_rows_written_here(reference_lmdb_dataset, tmp_path):
    ref, lmdb = reference_lmdb_dataset
    from sklearn.preprocessing import LabelEncoder
    label_encoders = {"instrument_family_str": LabelEncoder().fit(["bass", "brass", "flute"]),
                      "pitch": LabelEncoder().fit(list(range(24, 85)))}
    # the source protocol of extract.CodeExtractor / SpectrogramBatches: per-batch attribute columns
    attributes = {"instrument_family_str": torch.tensor([0, 2, 1]), "pitch": torch.tensor([36, 12, 60])}
    names = ["bass_synthetic_000-060-100", "flute_acoustic_002-036-050", "brass_acoustic_001-084-127"]
    rows = [extract.CodeRow(top=np.arange(128).reshape(32, 4) + i, bottom=np.arange(512).reshape(64, 8) * (i + 1),
                            attributes=extract._row_attributes(attributes, i), filename=n)
            for i, n in enumerate(names)]

    env = lmdb.open(str(tmp_path), map_size=1 << 28, max_dbs=2)
    codes_db = env.open_db("codes".encode("utf-8"), dupsort=False)          # extract_code.py:47-50
    extract.write_label_encoders(env, label_encoders)                        # extract_code.py:52-57
    assert extract.write_lmdb(rows, env, db=codes_db) == 3
    # the reference also dumps the encoders' classes next to the database (its reader wants the json)
    sys.modules[REF_MODULE.rsplit(".", 1)[0] + ".label_encoders"].dump_label_encoders(label_encoders, tmp_path)
    if hasattr(env, "close"):
        env.close()

    dataset = ref.LMDBDataset(tmp_path, classes_for_conditioning=["pitch", "instrument_family_str"])
    assert len(dataset) == 3
    order = sorted(range(3), key=lambda i: names[i].encode())                # LMDB keys are sorted
    for index, i in enumerate(order):
        top, bottom, attrs = dataset[index]
        assert top.dtype == torch.int64 and torch.equal(top, torch.from_numpy(rows[i].top))
        assert torch.equal(bottom, torch.from_numpy(rows[i].bottom))
        assert list(attrs) == ["pitch", "instrument_family_str"]
        assert attrs["pitch"].shape == (1,) and int(attrs["pitch"]) == int(attributes["pitch"][i])
        assert int(attrs["instrument_family_str"]) == int(attributes["instrument_family_str"][i])
    # the label encoders record of extract_code.py:55-57 unpickles to the encoders
    with lmdb.open(str(tmp_path), map_size=1 << 28, max_dbs=2).begin() as txn:
        stored = pickle.loads(txn.get("label_encoders".encode("utf-8")))
    assert list(stored["pitch"].classes_) == list(range(24, 85))
    assert type(pickle.loads(extract.row_record(rows[0])[1])) is ref.CodeRow


def test_extract_sources_carry_attributes_into_rows():
    """``_unpack_batch`` accepts both batch protocols; attribute values become the 0-dim tensors
    the reference stores (extract_code.py:70-75)."""
    spec = torch.zeros(2, 2, 8, 8)
    fam, pitch = torch.tensor([3, 1]), torch.tensor([40, 41])
    ours = (spec, ["a", "b"], {"instrument_family_str": fam, "pitch": pitch})
    theirs = (spec, fam, pitch, {"note_str": ["a", "b"], "categorical_fields": ["instrument_family_str", "pitch"]})
    for batch in (ours, theirs):
        s, names, attrs = extract._unpack_batch(batch)
        assert s is spec and list(names) == ["a", "b"]
        row1 = extract._row_attributes(attrs, 1)
        assert set(row1) == {"instrument_family_str", "pitch"}
        assert row1["pitch"].shape == () and int(row1["pitch"]) == 41 and row1["pitch"].dtype == torch.int64
    assert extract._unpack_batch((spec, ["a", "b"])) == (spec, ["a", "b"], None)
    assert extract._row_attributes(None, 0) == {}
