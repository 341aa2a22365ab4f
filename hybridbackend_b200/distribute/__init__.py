"""Mirror of hybridbackend.tensorflow.distribute for the hot path."""
from hybridbackend_b200.distribute.partition import partition_by_dual_modulo_stage_one
from hybridbackend_b200.distribute.partition import partition_by_dual_modulo_stage_two
from hybridbackend_b200.distribute.partition import partition_by_modulo
from hybridbackend_b200.distribute.collective import Collective
from hybridbackend_b200.distribute.collective import Topology
