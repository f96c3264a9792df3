import pytest

from dist_mnist_b200.cluster import ClusterSpec, default_device_index, parse_hosts, split_endpoint


def test_parse_hosts():
    assert parse_hosts("127.0.0.1:9900,127.0.0.1:9901", "worker_hosts") == ["127.0.0.1:9900", "127.0.0.1:9901"]
    with pytest.raises(ValueError):
        parse_hosts(None, "ps_hosts")
    with pytest.raises(ValueError):
        parse_hosts("localhost", "ps_hosts")
    assert split_endpoint("host-1:12") == ("host-1", 12)


def test_cluster_spec_readme_topology():
    c = ClusterSpec.from_flags("127.0.0.1:9910", "127.0.0.1:9900,127.0.0.1:9901")   # README.md:7-9
    assert c.num_ps == 1 and c.num_workers == 2 and c.num_tasks == 3
    assert c.as_dict() == {"worker": ["127.0.0.1:9900", "127.0.0.1:9901"], "ps": ["127.0.0.1:9910"]}
    assert c.task_endpoint("worker", 1) == "127.0.0.1:9901"
    assert c.rendezvous_endpoint() == ("127.0.0.1", 9910)
    with pytest.raises(ValueError):
        c.task_endpoint("worker", 2)
    with pytest.raises(ValueError):
        c.task_endpoint("chief", 0)


def test_device_mapping():
    c = ClusterSpec.from_flags("a:1,a:2", "a:3,a:4,a:5,a:6,a:7,a:8")   # 2 ps + 6 workers on 8 GPUs
    assert [default_device_index(c, "ps", k, 8) for k in range(2)] == [0, 1]
    assert [default_device_index(c, "worker", i, 8) for i in range(6)] == [2, 3, 4, 5, 6, 7]
    assert default_device_index(c, "worker", 0, 1) == 0          # everything on one GPU
    assert default_device_index(c, "worker", 1, 8, colocate=True) == 1
    assert default_device_index(c, "worker", 1, 0) == -1
