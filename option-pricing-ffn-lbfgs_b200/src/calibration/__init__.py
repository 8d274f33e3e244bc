# Calibration module
