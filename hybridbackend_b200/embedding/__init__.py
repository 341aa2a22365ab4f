"""Mirror of hybridbackend.tensorflow.embedding for the hot path."""
from hybridbackend_b200.embedding.sharding import is_small_table
from hybridbackend_b200.embedding.sharding import shard_offset
from hybridbackend_b200.embedding.sharding import shard_rows
from hybridbackend_b200.embedding.sharding import ShardedEmbeddingWeights
from hybridbackend_b200.embedding.lookup import embedding_lookup
from hybridbackend_b200.embedding.lookup import embedding_lookup_sparse
from hybridbackend_b200.embedding.lookup import GroupLookup
from hybridbackend_b200.embedding.lookup import segment_ids_to_offsets
from hybridbackend_b200.embedding.cache import lookup
from hybridbackend_b200.embedding.checkpoint import merge_shards
from hybridbackend_b200.embedding.checkpoint import logical_rows_of_merged
from hybridbackend_b200.embedding.checkpoint import split_merged
from hybridbackend_b200.embedding.transfer import h2d_transfer_n
