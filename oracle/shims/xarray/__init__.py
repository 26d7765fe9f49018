"""Minimal stand-in for the subset of xarray.DataArray the DIE reference's step path uses.
See oracle/shims/README.md.  NOT a general xarray replacement."""
import numpy as np
import pandas as pd

__version__ = "0.0-die-shim"


def _is_da(x):
    return isinstance(x, DataArray)


class _Coords:
    def __init__(self, owner):
        self._o = owner

    def __getitem__(self, name):
        return self._o._coord_array(name)

    def __contains__(self, name):
        return name in self._o._coords

    def keys(self):
        return self._o._coords.keys()


class _Loc:
    def __init__(self, owner):
        self._o = owner

    def __getitem__(self, key):
        return self._o._getitem_indexers(self._o._label_indexers(key))

    def __setitem__(self, key, value):
        self._o._setitem_indexers(self._o._label_indexers(key), value)


class DataArray:
    __array_priority__ = 60

    def __init__(self, data=None, coords=None, dims=None, name=None):
        if isinstance(data, DataArray):
            data = data.values
        coords = dict(coords) if coords is not None else {}
        if dims is None:
            arr = np.asarray(data)
            if coords and len(coords) == arr.ndim:
                dims = list(coords.keys())          # inferred from dict-like coords
            else:
                dims = [f"dim_{n}" for n in range(arr.ndim)]
        self.dims = tuple(dims)
        # coords: name -> (dims tuple, values); dimension coords are 1-D along their own dim
        self._coords = {}
        for k, v in coords.items():
            if isinstance(v, DataArray):
                self._coords[k] = (v.dims, np.asarray(v.values))
            elif isinstance(v, tuple) and len(v) == 2 and isinstance(v[0], (tuple, str)):
                d = (v[0],) if isinstance(v[0], str) else tuple(v[0])
                self._coords[k] = (d, np.asarray(v[1]))
            else:
                a = np.asarray(v)
                self._coords[k] = ((k,), a) if a.ndim == 1 else ((), a)
        arr = np.asarray(data)
        if arr.ndim == 0 and self.dims:               # scalar broadcast to the coords' shape
            shape = tuple(len(self._coords[d][1]) for d in self.dims)
            arr = np.full(shape, arr[()])
        self._data = np.array(arr) if not isinstance(arr, np.ndarray) else arr
        self.name = name

    # ---- basic protocol -------------------------------------------------------------------
    @property
    def values(self):
        return self._data

    def to_numpy(self):
        return self._data

    def __array__(self, dtype=None, copy=None):
        return self._data if dtype is None else self._data.astype(dtype)

    @property
    def shape(self):
        return self._data.shape

    @property
    def ndim(self):
        return self._data.ndim

    @property
    def dtype(self):
        return self._data.dtype

    def __len__(self):
        return self._data.shape[0]

    @property
    def coords(self):
        return _Coords(self)

    @property
    def loc(self):
        return _Loc(self)

    def copy(self, deep=True):
        out = self._like(self._data.copy() if deep else self._data)
        return out

    def _like(self, data, dims=None, coords=None):
        out = DataArray.__new__(DataArray)
        out.dims = self.dims if dims is None else tuple(dims)
        out._coords = dict(self._coords) if coords is None else coords
        out._data = data
        out.name = self.name
        return out

    def _coord_array(self, name):
        d, v = self._coords[name]
        return DataArray(v, dims=d, coords={name: (d, v)} if len(d) == 1 and d[0] == name else {})

    def __iter__(self):
        for k in range(self._data.shape[0]):
            yield self.isel({self.dims[0]: k})

    def __repr__(self):
        return f"<shim DataArray dims={self.dims} shape={self.shape}>"

    # ---- indexing -------------------------------------------------------------------------
    def _index_of(self, dim):
        return pd.Index(self._coords[dim][1])

    def _label_indexers(self, key, method=None):
        """labels -> positional indexers (ints, int arrays or DataArrays of ints)."""
        out = {}
        for dim, lab in key.items():
            index = self._index_of(dim)
            if isinstance(lab, DataArray):
                pos = index.get_indexer(np.asarray(lab.values).ravel(), method=method).reshape(lab.shape)
                if (pos < 0).any():
                    raise KeyError(f"labels not found along {dim!r}")
                out[dim] = DataArray(pos, dims=lab.dims)
            elif isinstance(lab, (list, tuple, np.ndarray)):
                pos = index.get_indexer(np.asarray(lab), method=method)
                if (pos < 0).any():
                    raise KeyError(f"labels {lab!r} not found along {dim!r}")
                out[dim] = pos
            else:
                pos = index.get_indexer([lab], method=method)[0]
                if pos < 0:
                    raise KeyError(f"label {lab!r} not found along {dim!r}")
                out[dim] = int(pos)
        return out

    def _np_key(self, indexers):
        """positional indexers -> (numpy key, new dims, kept-coords builder).  DataArray indexers
        that share a dimension are applied POINTWISE (xarray's vectorised indexing); plain
        lists / arrays index orthogonally along their own dimension."""
        vec_dims = []
        for v in indexers.values():
            if isinstance(v, DataArray):
                for d in v.dims:
                    if d not in vec_dims:
                        vec_dims.append(d)
        if len(vec_dims) > 1:
            raise NotImplementedError("shim: vectorised indexers must share one dimension")
        n_list = sum(isinstance(v, (list, np.ndarray)) for v in indexers.values())
        if vec_dims and n_list:
            raise NotImplementedError("shim: mixing vectorised and orthogonal indexers")
        key, new_dims, placed_vec = [], [], False
        for d in self.dims:
            v = indexers.get(d, slice(None))
            if isinstance(v, DataArray):
                key.append(np.asarray(v.values))
                if not placed_vec:
                    new_dims.append(vec_dims[0])
                    placed_vec = True
            elif isinstance(v, (int, np.integer)):
                key.append(int(v))
            elif isinstance(v, slice):
                key.append(v)
                new_dims.append(d)
            else:
                key.append(np.asarray(v))
                new_dims.append(d)
        return tuple(key), new_dims, vec_dims

    def _fix_vectorised_axis_order(self, key, new_dims, vec_dims):
        """numpy puts the broadcast (advanced) axis first when advanced indices are separated by
        slices; xarray keeps the position of the first indexed dim.  Returns a transpose or None."""
        adv = [i for i, k in enumerate(key) if isinstance(k, np.ndarray)]
        if not vec_dims or len(adv) < 2:
            return None
        contiguous = adv == list(range(adv[0], adv[-1] + 1))
        if contiguous:
            return None
        # numpy result: (vec, <sliced dims...>); wanted: new_dims order
        np_dims = [vec_dims[0]] + [d for d in new_dims if d != vec_dims[0]]
        return [np_dims.index(d) for d in new_dims]

    def _getitem_indexers(self, indexers):
        key, new_dims, vec_dims = self._np_key(indexers)
        data = self._data[key]
        perm = self._fix_vectorised_axis_order(key, new_dims, vec_dims)
        if perm is not None:
            data = data.transpose(perm)
        coords = {}
        for name, (d, v) in self._coords.items():
            if all(dd in self.dims for dd in d) and len(d) == 1:
                dim = d[0]
                ind = indexers.get(dim, slice(None))
                if isinstance(ind, DataArray):
                    coords[name] = (tuple(vec_dims), v[np.asarray(ind.values)])
                elif isinstance(ind, (int, np.integer)):
                    coords[name] = ((), np.asarray(v[int(ind)]))
                elif isinstance(ind, slice):
                    coords[name] = (d, v[ind])
                else:
                    coords[name] = (d, v[np.asarray(ind)])
            else:
                coords[name] = (d, v)
        return self._like(np.asarray(data), dims=new_dims, coords=coords)

    def _setitem_indexers(self, indexers, value):
        key, new_dims, vec_dims = self._np_key(indexers)
        val = value.values if isinstance(value, DataArray) else value
        perm = self._fix_vectorised_axis_order(key, new_dims, vec_dims)
        if perm is not None and np.ndim(val) == len(new_dims):
            inv = np.argsort(perm)
            val = np.asarray(val).transpose(inv)
        self._data[key] = val                   # numpy fancy assignment: last writer wins

    def isel(self, indexers=None, **kw):
        indexers = dict(indexers or {}, **kw)
        norm = {}
        for d, v in indexers.items():
            norm[d] = v
        return self._getitem_indexers(norm)

    def sel(self, indexers=None, method=None, **kw):
        key = dict(indexers or {}, **kw)
        return self._getitem_indexers(self._label_indexers(key, method=method))

    def __getitem__(self, key):
        if isinstance(key, str):
            return self._coord_array(key)
        if isinstance(key, dict):
            return self.isel(key)
        if not isinstance(key, tuple):
            key = (key,)
        indexers = {}
        for d, k in zip(self.dims, key):
            indexers[d] = k
        return self._getitem_indexers(indexers)

    def __setitem__(self, key, value):
        if isinstance(key, dict):
            self._setitem_indexers(key, value)
            return
        val = value.values if isinstance(value, DataArray) else value
        self._data[key] = val

    # ---- arithmetic (operands aligned by dimension NAME) -------------------------------------
    def _aligned(self, other):
        if not isinstance(other, DataArray):
            return self.dims, self._data, other
        dims = list(self.dims) + [d for d in other.dims if d not in self.dims]

        def expand(a):
            data = a._data
            order = [d for d in dims if d in a.dims]
            data = data.transpose([a.dims.index(d) for d in order]) if order != list(a.dims) else data
            shape = [data.shape[order.index(d)] if d in order else 1 for d in dims]
            return data.reshape(shape)

        return tuple(dims), expand(self), expand(other)

    def _binary(self, other, op, reflexive=False):
        dims, a, b = self._aligned(other)
        res = op(b, a) if reflexive else op(a, b)
        coords = dict(self._coords)
        if isinstance(other, DataArray):
            for k, v in other._coords.items():
                coords.setdefault(k, v)
        return self._like(np.asarray(res), dims=dims, coords=coords)

    def __add__(self, o): return self._binary(o, np.add)
    def __radd__(self, o): return self._binary(o, np.add, True)
    def __sub__(self, o): return self._binary(o, np.subtract)
    def __rsub__(self, o): return self._binary(o, np.subtract, True)
    def __mul__(self, o): return self._binary(o, np.multiply)
    def __rmul__(self, o): return self._binary(o, np.multiply, True)
    def __truediv__(self, o): return self._binary(o, np.true_divide)
    def __rtruediv__(self, o): return self._binary(o, np.true_divide, True)
    def __mod__(self, o): return self._binary(o, np.remainder)
    def __gt__(self, o): return self._binary(o, np.greater)
    def __ge__(self, o): return self._binary(o, np.greater_equal)
    def __lt__(self, o): return self._binary(o, np.less)
    def __le__(self, o): return self._binary(o, np.less_equal)
    def __and__(self, o): return self._binary(o, np.logical_and)
    def __or__(self, o): return self._binary(o, np.logical_or)
    def __neg__(self): return self._like(-self._data)
    def __abs__(self): return self._like(np.abs(self._data))

    def _inplace(self, other, op):
        dims, a, b = self._aligned(other)
        if tuple(dims) != self.dims:
            raise ValueError("shim: in-place operand would add dimensions")
        self._data = op(a, b)
        return self

    def __iadd__(self, o): return self._inplace(o, np.add)
    def __isub__(self, o): return self._inplace(o, np.subtract)
    def __imul__(self, o): return self._inplace(o, np.multiply)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__" or kwargs.get("out") is not None:
            return NotImplemented
        das = [x for x in inputs if isinstance(x, DataArray)]
        if len(inputs) == 1:
            return das[0]._like(np.asarray(ufunc(das[0]._data, **kwargs)))
        if len(inputs) == 2:
            if isinstance(inputs[0], DataArray):
                return inputs[0]._binary(inputs[1], lambda a, b: ufunc(a, b, **kwargs))
            return inputs[1]._binary(inputs[0], lambda a, b: ufunc(a, b, **kwargs), reflexive=True)
        return NotImplemented

    # ---- reductions / masking ----------------------------------------------------------------
    def sum(self, dim=None, skipna=None):
        if dim is not None:
            raise NotImplementedError
        f = np.nansum if np.issubdtype(self._data.dtype, np.floating) and skipna is not False else np.sum
        return DataArray(f(self._data))

    def where(self, cond, other=np.nan):
        dims, a, c = self._aligned(cond)
        if tuple(dims) != self.dims:
            raise ValueError("shim: where() condition adds dimensions")
        return self._like(np.where(np.asarray(c).astype(bool), a, other))

    def dropna(self, dim, how="any"):
        ax = self.dims.index(dim)
        other_axes = tuple(i for i in range(self.ndim) if i != ax)
        bad = np.isnan(self._data).any(axis=other_axes) if how == "any" else np.isnan(self._data).all(axis=other_axes)
        keep = np.nonzero(~bad)[0]
        return self.isel({dim: keep})

    def clip(self, a_min=None, a_max=None):
        return self._like(np.clip(self._data, a_min, a_max))

    def round(self, decimals=0):
        return self._like(np.round(self._data, decimals))

    def __float__(self):
        return float(self._data)

    def __bool__(self):
        return bool(self._data)
