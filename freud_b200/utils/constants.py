"""Constants of src/utils/constants.py that the search path needs (reference :6,17)."""
SAMPLE_RATE = 16000
TIMESTEP_S = 30 / 1500
