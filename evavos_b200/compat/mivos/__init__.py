"""Drop-in `mivos` package backed by evavos_b200 (see ../README.md)."""
