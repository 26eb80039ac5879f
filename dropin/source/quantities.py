"""Drop-in replacement for HELIOS's source/quantities.py: `Store` with the reference's public surface
(quantities.py:29-665), its `dev_*` handles owned by libhelios_b200.so.  Copy (or symlink) this file over
source/quantities.py of a HELIOS checkout; see INTEGRATION.md."""
from helios_b200.quantities import Store  # noqa: F401

if __name__ == "__main__":
    print("This module is for storing and allocating all the necessary quantities used in HELIOS (B200 backend).")
