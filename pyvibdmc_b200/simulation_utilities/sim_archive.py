"""Checkpoints (pickle, protocol 4) and HDF5 dumps with the reference's file names and keys
(simulation_utilities/sim_archive.py:13-43).  h5py is used when importable, otherwise the in-tree
writer (h5lite), which emits the same bytes h5py does for these files."""
import copy
import glob
import pickle

try:                                    # pragma: no cover - h5py is absent from the B200 image
    import h5py as _h5
except ImportError:                     # the normal case here
    from . import h5lite as _h5

__all__ = ['SimArchivist']


class SimArchivist:
    @staticmethod
    def save_h5(fname, keyz, valz):
        with _h5.File(fname, 'w') as hf:
            for key, val in zip(keyz, valz):
                hf.create_dataset(key, data=val)

    @staticmethod
    def chkpt(dmc_obj, prop_step):
        snapshot = copy.deepcopy(dmc_obj)          # DMC_Sim.__deepcopy__ drops the un-picklable plug-ins
        with open(f'{dmc_obj.output_folder}/chkpts/{dmc_obj.sim_name}_{str(prop_step)}.pickle', 'wb') as handle:
            pickle.dump(snapshot, handle, protocol=4)

    @staticmethod
    def reload_sim(chkpt_folder, sim_name):
        pickle_file = glob.glob(f"{chkpt_folder}/chkpts/{sim_name}_*.pickle")[0]
        with open(pickle_file, "rb") as handle:
            return pickle.load(handle)
