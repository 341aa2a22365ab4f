import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA GPU (run on the B200 box)')


@pytest.fixture(scope='session')
def oracle():
  from oracle import hb_oracle
  hb_oracle.lib()
  return hb_oracle


@pytest.fixture(scope='session')
def golden_partition():
  return dict(np.load(os.path.join(GOLDEN, 'partition_ref.npz')))


@pytest.fixture(scope='session')
def golden_alltoall():
  import json
  with open(os.path.join(GOLDEN, 'alltoall_ref.json')) as f:
    return json.load(f)


@pytest.fixture(scope='session')
def hb():
  import hybridbackend_b200
  hybridbackend_b200._lib.lib()
  return hybridbackend_b200


def criteo_table_sizes():
  # docs/tutorial/ranking/criteo/data/spec.json:105-355 (Criteo-Terabyte vocabularies)
  return [39884406, 39043, 17289, 7420, 20263, 3, 7120, 1543, 63, 38532951, 2953546,
          403346, 10, 2208, 11938, 155, 4, 976, 14, 39979771, 25641295, 39664984,
          585935, 12972, 108, 36]
