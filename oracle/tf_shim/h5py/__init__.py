"""Import-only stand-in for h5py (absent from the image): lets the reference's dataset.py be imported so that its batch
assembly (split_entries / trim_collate, dataset.py:288-355) can be executed on in-memory arrays by
oracle/make_golden_ref_collate.py.  Opening a file is not supported.  TEST INFRASTRUCTURE."""


class File:
    def __init__(self, *a, **k):
        raise NotImplementedError("h5py stand-in: files cannot be opened (the real h5py is not installed)")
