"""Constants shared with the reference (treeCl/constants.py:9-10,21)."""
import re

SORT_KEY = lambda item: tuple((int(num) if num else alpha) for (num, alpha) in
                              re.findall(r'(\d+)|(\D+)', item))
RANDOM_SEED = int('CA55E77E', 16)
