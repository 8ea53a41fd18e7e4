"""Three more episodes of the reference's result files followed by the CUDA path (see test_golden_gpu.py).  They were added with
the 200-episode pin after the last GPU session of round 1 -- the oracle follows them for 55 / 60 / 60 rows
(tests/golden/oracle_golden_scan.json) -- so their lower bound stays at 10 rows until the CUDA path's own count has been read
off a GPU run (the test prints it), and the file sorts after the suites that have been seen green on a B200."""
import pytest

from test_golden_gpu import follow_episode

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("map_size,seed,n_steps,n_min", [(40, 12, 50, 10), (60, 7, 55, 10), (100, 6, 55, 10)])
def test_cuda_path_tracks_more_reference_episodes(map_size, seed, n_steps, n_min):
    follow_episode(map_size, seed, n_steps, n_min)
