"""Drop-in mirror of the reference package layout (src/models, src/calibration, src/data)."""
