"""Inter-stage data-plane transports."""
from .base import ChaosPolicy, Message, MessageQueue, Transport, TransportError, build_msg  # noqa: F401
from .inproc import LoopbackTransport, QueueTransport, ring  # noqa: F401
from .socket_transport import InputNodeConnection, OutputNodeConnection, SocketTransport  # noqa: F401
