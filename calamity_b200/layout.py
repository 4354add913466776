"""Ragged description of the foreground basis handed to the native library.

The reference stores the basis as dense zero-padded tensors, one per chunk
(`tensorize_fg_model_comps_dict`, /root/reference/calamity/calibration.py:104-190).  The native
library never materialises the padding: a fitting group is (ncomp coefficients) x (nslots redundant
sub-groups, each repeated for its baselines).  This module converts both the reference's dict format
and its dense tensors into that description, and maps flat device vectors back to the reference's
per-chunk tensor shapes.  Orderings are the reference's: chunk -> group -> redundant group -> baseline.
"""
import numpy as np


class RaggedLayout:
    def __init__(self, nants, nfreqs, dtype=np.float32):
        self.nants = int(nants)
        self.nfreqs = int(nfreqs)
        self.dtype = np.dtype(dtype)  # element type of the basis blocks and of the plan built from this layout
        self.group_ncomp = []
        self.group_nslots = []
        self.slot_nbls = []
        self.bl_ant0 = []
        self.bl_ant1 = []
        self.blocks = []          # per group float32 [nslots, ncomp, nfreqs]; objects may be shared
        self.chunks = []          # per chunk dict(nvecs, ngrps, nbls, group0)
        self._finalized = False

    # ------------------------------------------------------------------ construction
    def _add_group(self, block, slot_baselines):
        nslots, ncomp, nf = block.shape
        assert nf == self.nfreqs and nslots == len(slot_baselines)
        self.group_ncomp.append(ncomp)
        self.group_nslots.append(nslots)
        for bls in slot_baselines:
            self.slot_nbls.append(len(bls))
            for (i, j) in bls:
                self.bl_ant0.append(i)
                self.bl_ant1.append(j)
        self.blocks.append(block)

    def _finalize(self):
        self.group_ncomp = np.asarray(self.group_ncomp, dtype=np.int32)
        self.group_nslots = np.asarray(self.group_nslots, dtype=np.int32)
        self.slot_nbls = np.asarray(self.slot_nbls, dtype=np.int32)
        self.bl_ant0 = np.asarray(self.bl_ant0, dtype=np.int32)
        self.bl_ant1 = np.asarray(self.bl_ant1, dtype=np.int32)
        self.group_coef0 = np.concatenate([[0], np.cumsum(self.group_ncomp)]).astype(np.int64)
        self.group_class = self._basis_classes()
        self.ngroups = len(self.group_ncomp)
        self.ncoef = int(self.group_coef0[-1])
        self.nbls = len(self.bl_ant0)
        self._finalized = True
        return self

    def _basis_classes(self):
        """int32 [ngroups]: equal ids <=> identical basis blocks.  modeling.yield_pbl_dpss_model_comps gives every baseline
        of one integer-ns delay the SAME ndarray (/root/reference/calamity/modeling.py:293, operator cache 352/371), so
        object identity finds the classes for free; distinct objects are additionally merged by a digest of their bytes
        (a caller that deep-copied the dict still gets one class per distinct basis)."""
        import hashlib

        by_obj, by_digest = {}, {}
        out = np.empty(len(self.blocks), dtype=np.int32)
        for g, blk in enumerate(self.blocks):
            cid = by_obj.get(id(blk))
            if cid is None:
                a = np.ascontiguousarray(blk)
                key = (a.shape, a.dtype.str, hashlib.blake2b(a.view(np.uint8).reshape(-1), digest_size=16).digest())
                cid = by_digest.setdefault(key, len(by_digest))
                by_obj[id(blk)] = cid
            out[g] = cid
        return out

    @classmethod
    def from_chunked_dict(cls, chunked, ants_map, nfreqs, nants=None, dtype=np.float32):
        """`chunked` is the output of chunk_fg_comp_dict_by_nbls: {(nbl, nvecs): {fit_grp: [nrg*nfreqs, ncomp]}}."""
        lay = cls(len(ants_map) if nants is None else nants, nfreqs, dtype=dtype)
        cache = {}
        for (nbls, nvecs), grp_dict in chunked.items():
            lay.chunks.append(dict(nvecs=int(nvecs), ngrps=len(grp_dict), nbls=int(nbls), group0=len(lay.blocks)))
            for fit_grp, vecs in grp_dict.items():
                key = id(vecs)
                if key not in cache:
                    nrg = len(fit_grp)
                    blk = np.asarray(vecs).reshape(nrg, nfreqs, vecs.shape[1]).transpose(0, 2, 1)
                    cache[key] = (np.ascontiguousarray(blk, dtype=lay.dtype), vecs)  # keep vecs alive: id() is reused otherwise
                slots = [[(ants_map[ap[0]], ants_map[ap[1]]) for ap in red] for red in fit_grp]
                lay._add_group(cache[key][0], slots)
        return lay._finalize()

    @classmethod
    def from_dense(cls, fg_comps, corr_inds, nants, dtype=np.float32):
        """Reference tensors [nvecs, ngrps, nbls, nfreqs] + corr_inds.  Trailing all-zero rows are dropped
        (they never change the model nor receive gradient); every baseline becomes its own slot."""
        nfreqs = int(np.shape(fg_comps[0])[3])
        lay = cls(nants, nfreqs, dtype=dtype)
        for comps, chunk in zip(fg_comps, corr_inds):
            comps = np.asarray(comps)
            nvecs, ngrps, nbls, _ = comps.shape
            lay.chunks.append(dict(nvecs=nvecs, ngrps=ngrps, nbls=nbls, group0=len(lay.blocks)))
            for g in range(ngrps):
                rows = comps[:, g]
                live = np.where(np.any(rows.reshape(nvecs, -1) != 0, axis=1))[0]
                ncomp = int(live.max()) + 1 if len(live) else 0
                blk = np.ascontiguousarray(rows[:ncomp].transpose(1, 0, 2), dtype=lay.dtype)
                lay._add_group(blk, [[(int(i), int(j))] for (i, j) in chunk[g]])
        return lay._finalize()

    # ------------------------------------------------------------------ reference-shape <-> flat
    def flatten_data(self, chunk_tensors):
        """list of [ngrps, nbls, nfreqs] -> [nbls_total, nfreqs] (layout dtype) in canonical baseline order."""
        return np.ascontiguousarray(
            np.concatenate([np.asarray(t, dtype=self.dtype).reshape(-1, self.nfreqs) for t in chunk_tensors], axis=0)
        )

    def unflatten_data(self, flat, dtype=np.float32):
        out, b = [], 0
        for ch in self.chunks:
            n = ch["ngrps"] * ch["nbls"]
            out.append(np.asarray(flat[b : b + n], dtype=dtype).reshape(ch["ngrps"], ch["nbls"], self.nfreqs))
            b += n
        return out

    def chunk_coef_bounds(self):
        """[nchunks + 1] coefficient offsets of the chunks: chunk c is the reference's variable pair fg_r[c] / fg_i[c]
        (calibration.py:560-567).  Chunks without stored coefficients are skipped."""
        b = [0]
        for ch in self.chunks:
            e = int(self.group_coef0[ch["group0"] + ch["ngrps"]])
            if e > b[-1]:
                b.append(e)
        return np.asarray(b, dtype=np.int64)

    def flatten_coeffs(self, chunk_coeffs):
        """list of [nvecs, ngrps, 1, 1] -> [ncoef] (layout dtype)."""
        flat = np.zeros(self.ncoef, dtype=self.dtype)
        for ch, t in zip(self.chunks, chunk_coeffs):
            t = np.asarray(t).reshape(ch["nvecs"], ch["ngrps"])
            for g in range(ch["ngrps"]):
                gg = ch["group0"] + g
                n = self.group_ncomp[gg]
                flat[self.group_coef0[gg] : self.group_coef0[gg] + n] = t[:n, g]
        return flat

    def unflatten_coeffs(self, flat, template=None, dtype=np.float32):
        """[ncoef] -> list of [nvecs, ngrps, 1, 1]; rows above a group's ncomp keep the template's values
        (zero padding stays exactly as given: zero basis rows get zero gradient, quirk Q11)."""
        out = []
        for c, ch in enumerate(self.chunks):
            if template is not None:
                t = np.array(np.asarray(template[c]).reshape(ch["nvecs"], ch["ngrps"]), dtype=dtype)
            else:
                t = np.zeros((ch["nvecs"], ch["ngrps"]), dtype=dtype)
            for g in range(ch["ngrps"]):
                gg = ch["group0"] + g
                n = self.group_ncomp[gg]
                t[:n, g] = flat[self.group_coef0[gg] : self.group_coef0[gg] + n]
            out.append(t.reshape(ch["nvecs"], ch["ngrps"], 1, 1))
        return out

    def corr_inds(self):
        """The reference's corr_inds[chunk][group][baseline] = (i, j) list structure."""
        out, b, s = [], 0, 0
        for ch in self.chunks:
            chunk = []
            for g in range(ch["ngrps"]):
                gg = ch["group0"] + g
                grp = []
                for _ in range(self.group_nslots[gg]):
                    for _ in range(self.slot_nbls[s]):
                        grp.append((int(self.bl_ant0[b]), int(self.bl_ant1[b])))
                        b += 1
                    s += 1
                chunk.append(grp)
            out.append(chunk)
        return out

    def dense_chunks(self, dtype=np.float32):
        """The reference's zero-padded [nvecs, ngrps, nbls, nfreqs] tensors (for API compatibility / tests)."""
        out, s = [], 0
        for ch in self.chunks:
            dense = np.zeros((ch["nvecs"], ch["ngrps"], ch["nbls"], self.nfreqs), dtype=dtype)
            for g in range(ch["ngrps"]):
                gg = ch["group0"] + g
                blk = self.blocks[gg]
                b = 0
                for sl in range(self.group_nslots[gg]):
                    for _ in range(self.slot_nbls[s]):
                        dense[: blk.shape[1], g, b] = blk[sl]
                        b += 1
                    s += 1
            out.append(dense)
        return out

    # ------------------------------------------------------------------ sizes
    def sizes(self):
        n_a_nz = int(np.sum(np.repeat(self.group_ncomp, 1) * self.group_nslots)) * self.nfreqs
        return dict(
            n_d=self.nbls * self.nfreqs,
            n_a_nz=n_a_nz,
            n_c_nz=self.ncoef,
            b_iter=4 * n_a_nz + 12 * self.nbls * self.nfreqs + 48 * (self.ncoef + self.nants * self.nfreqs),
        )
