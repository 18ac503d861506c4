"""Exceptions of the batched engine."""


class NotLowerableError(TypeError):
    """The env uses an agent / payload / resolver feature that has no device program.

    phantom_b200 never falls back to executing Python handlers on the CPU: an env either
    lowers to a fused CUDA kernel or construction fails with this error."""


class DeviceOnlyError(RuntimeError):
    """A method that the reference executes in Python per message (Network.send,
    Agent.handle_message, ...) was called on the host.  In phantom_b200 that work happens
    inside the fused step kernel."""
