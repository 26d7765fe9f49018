"""Fast-path hints between an Env and the agents acting on its observation.

``Env.step`` caches, per slot, the linear cell index of every agent, and -- once a gradient agent has
asked for it -- publishes np.gradient of the new chem1 channel per cell.  ``Agent.forward(obs)`` only
receives ``obs``; it finds the producing Env here (``find_env``), keyed by the medium tensor's device address,
and ``Env._forward_flags`` lets it use the cached arrays ONLY IF the observation is provably the one the Env
wrote in its last step:
same storage, and torch's in-place version counters of the medium / agents tensors unchanged (our
kernels write through raw pointers and do not bump them; any ``tensor[...] = ...`` by the caller
does).  Otherwise the kernel recomputes everything from ``obs`` -- same results bit for bit."""
import weakref

_REGISTRY = {}          # medium data_ptr -> weakref(Env)


def publish(env, medium_buf) -> None:
    if len(_REGISTRY) > 4096:               # entries of collected Envs are dead weakrefs: drop them now and then
        for key in [k for k, r in _REGISTRY.items() if r() is None]:
            del _REGISTRY[key]
    _REGISTRY[(medium_buf.device.index, medium_buf.data_ptr())] = weakref.ref(env)


def find_env(agents, medium):
    """The live Env one of whose medium buffers `medium` is (same storage, same shape, and an agents tensor
    of that Env's size), or None.  Which of the Env's caches are valid for this observation is decided by
    ``Env._forward_flags``."""
    ref = _REGISTRY.get((medium.device.index, medium.data_ptr()))
    env = ref() if ref is not None else None
    if env is None or env._handle is None:
        return None
    buf = env._medium_buf[0]
    if medium.numel() != buf.numel() or medium.shape[-2:] != buf.shape[-2:] or agents.numel() != env._agents.numel() \
            or medium.dtype != buf.dtype:
        return None
    return env


# ---- host-buffer path: the action array Agent.forward returned, and its device copy -------------------------------
# Agent.forward(host obs) returns a READ-ONLY numpy view of a pinned buffer whose content was just downloaded from the
# agent's device action tensor.  If Env.step receives that very array (same address, still read-only) and the device
# tensor has not been written since (torch version counter), the step reads the device copy: no H2D of the action.
# A caller who wants to edit the action has to copy it (the array is read-only) -- and a copy takes the upload path.
_HOST_ACTIONS = {}      # host address -> (weakref(device tensor), version, nbytes)


def register_host_action(host_array, device_tensor) -> None:
    if len(_HOST_ACTIONS) > 256:
        for key in [k for k, (r, _, _) in _HOST_ACTIONS.items() if r() is None]:
            del _HOST_ACTIONS[key]
    _HOST_ACTIONS[host_array.ctypes.data] = (weakref.ref(device_tensor), device_tensor._version, host_array.nbytes)


def device_copy_of(host_array, device):
    """The device tensor `host_array` was downloaded from, if that is provably still its content; else None."""
    if host_array.flags.writeable:
        return None
    entry = _HOST_ACTIONS.get(host_array.ctypes.data)
    if entry is None:
        return None
    ref, version, nbytes = entry
    t = ref()
    if t is None or t._version != version or nbytes != host_array.nbytes or t.device != device:
        return None
    return t
