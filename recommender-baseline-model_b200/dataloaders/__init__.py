"""Host-side producers of the hot path's input wire formats + synthetic interaction data in
the ``[user_train, user_valid, user_test, usernum, itemnum]`` layout of ``data_partition``
(NN/dataloaders/__init__.py:17-66).  Text parsing, pickle caches and the popularity negative sampler of the reference
are out of scope.  ``device`` holds the on-device batch construction."""
from .synthetic import (synthetic_interactions, sliding_window_partition, BertBatcher, SasBatcher, eval_sequences,  # noqa: F401
                        uniform_negative_candidates)
from .device import (histories_to_csr, DeviceBertTrainLoader, DeviceSasTrainLoader, seen_sets_to_csr, popularity_cdf,  # noqa: F401,E402
                     DeviceNegativeSampler, DeviceEvalLoader)
from .partition import load_interactions_text, data_partition  # noqa: F401,E402
from .factory import DATALOADERS, dataloader_factory, BertDataloader, SASDataLoader  # noqa: F401,E402
