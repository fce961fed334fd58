"""Inter-stage data-plane transports."""
from .base import ChaosPolicy, Message, MessageQueue, Transport, TransportError, build_msg  # noqa: F401
from .inproc import LoopbackTransport, QueueTransport, ring  # noqa: F401
from .socket_transport import InputNodeConnection, OutputNodeConnection, SocketTransport  # noqa: F401


def __getattr__(name):  # lazy: importing torch.distributed is not needed for the socket path
    if name in ("TorchDistTransport", "make_edge_groups"):
        from . import nccl_p2p

        return getattr(nccl_p2p, name)
    raise AttributeError(name)
