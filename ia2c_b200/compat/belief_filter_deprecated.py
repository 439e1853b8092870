"""Drop-in for the module ``ia2c.py`` imports as ``belief_filter`` (ia2c.py:23; the reference ships it
only as belief_filter_deprecated.py)."""
import numpy as np  # noqa: F401

from ia2c_b200.belief import BeliefFilter, generate_random_probability_matrix  # noqa: F401
