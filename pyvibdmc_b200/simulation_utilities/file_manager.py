"""Output-folder layout and checkpoint garbage collection (reference: simulation_utilities/file_manager.py)."""
import glob
import os

__all__ = ['FileManager']


def _step_of(path, suffix_sep):
    return int(os.path.basename(path).split('_')[-1].split(suffix_sep)[0])


class FileManager:
    @staticmethod
    def create_filesystem(output_folder):
        """output_folder/{chkpts,wfns} (file_manager.py:41-54)."""
        for sub in ('', 'chkpts', 'wfns'):
            os.makedirs(os.path.join(output_folder, sub), exist_ok=True)

    @staticmethod
    def delete_older_checkpoints(sim_folder, sim_name, time_step):
        """Remove pickles written before `time_step` (file_manager.py:29-38)."""
        for p in sorted(glob.glob(f'{sim_folder}/chkpts/{sim_name}*.pickle')):
            if _step_of(p, '.') < time_step:
                os.remove(p)

    @staticmethod
    def delete_future_checkpoints(sim_folder, sim_name, time_step):
        """When restarting: drop checkpoints and wave functions later than `time_step` (file_manager.py:11-26)."""
        for p in sorted(glob.glob(f'{sim_folder}/chkpts/{sim_name}_*.pickle')):
            if _step_of(p, '.') > time_step:
                os.remove(p)
        for w in sorted(glob.glob(f'{sim_folder}/wfns/{sim_name}_*.hdf5')):
            if _step_of(w, 'ts') > time_step:
                os.remove(w)
