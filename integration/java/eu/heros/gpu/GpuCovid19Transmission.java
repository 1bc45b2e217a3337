package eu.heros.gpu;

import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;

import eu.heros.disease.Covid19Progression;
import gnu.trove.set.TIntSet;
import nl.tudelft.simulation.medlabs.common.MedlabsRuntimeException;
import nl.tudelft.simulation.medlabs.disease.DiseaseTransmission;
import nl.tudelft.simulation.medlabs.disease.InfectionRecord;
import nl.tudelft.simulation.medlabs.location.Location;
import nl.tudelft.simulation.medlabs.model.MedlabsModelInterface;

/**
 * Drop-in for eu.heros.disease.Covid19TransmissionArea and Covid19TransmissionDistance: the same constructor contract
 * (reads covidT_area.* / covidT_dist.* from the model's parameter map), the same two overrides. Registered in
 * ConstructHerosModel.java:136-144 in place of the two reference classes:
 *
 * <pre>
 * HerosGpu gpu = new HerosGpu(area ? HerosGpu.MODEL_AREA : HerosGpu.MODEL_DISTANCE, HerosGpu.TRIGGER_EXIT_ONLY, 0, seed, 24.0);
 * DiseaseProgression covidProgression = new GpuCovid19Progression(this.model, gpu);
 * DiseaseTransmission covidTransmission = new GpuCovid19Transmission(this.model, gpu, area);
 * </pre>
 *
 * Two modes. {@link #infectPeople} forwards one call to heros_infect_people (the 1:1 shim: same InfectionRecord as the
 * reference method would build, with keyed Philox draws in place of model.getU01()); this is what a parity run uses.
 * In production mode ({@link #setWindowed}) the calls of the medlabs engine return an uncalculated record -- nothing
 * is applied by medlabs -- while the finished stays are collected through {@link #onStayClosed} and carried out in
 * one batch per window by {@link GpuWindow}. Not compiled here (see {@link HerosGpu}).
 */
public class GpuCovid19Transmission extends DiseaseTransmission
{
    /** */
    private static final long serialVersionUID = 1L;

    private final HerosGpu gpu;

    private final boolean area;

    /** Production mode: stays are batched, infectPeople does nothing. */
    private boolean windowed = false;

    /** The stays that ended in the current window. */
    final VisitBuffer visits = new VisitBuffer(1 << 20);

    /**
     * @param model MedlabsModelInterface; the Medlabs model
     * @param gpu HerosGpu; the engine of this replication
     * @param area boolean; area-based (covidT_area.*) or distance-based (covidT_dist.*) transmission
     */
    public GpuCovid19Transmission(final MedlabsModelInterface model, final HerosGpu gpu, final boolean area)
    {
        super(model, "Covid19");
        this.gpu = gpu;
        this.area = area;
        // raw property values: the library converts days -> hours and seconds -> hours exactly as
        // Covid19TransmissionArea.java:63-68 / Covid19TransmissionDistance.java:59-68 do
        try (Arena a = Arena.ofConfined())
        {
            if (area)
            {
                String[] names = {"contagiousness", "beta", "t_e_min", "t_e_mode", "t_e_max", "calculation_threshold"};
                MemorySegment p = a.allocate(HerosGpuNative.HEROS_AREA_PARAMS_LAYOUT);
                for (int i = 0; i < names.length; i++)
                    p.setAtIndex(JAVA_DOUBLE, i, model.getParameterValueDouble("covidT_area." + names[i]));
                gpu.check((int) HerosGpuNative.SET_TRANSMISSION_AREA.invokeExact(gpu.handle, p));
            }
            else
            {
                String[] names = {"L", "I", "C", "v_max", "v_0", "r", "psi", "alpha", "mu", "calculation_threshold"};
                MemorySegment p = a.allocate(HerosGpuNative.HEROS_DIST_PARAMS_LAYOUT);
                for (int i = 0; i < names.length; i++)
                    p.setAtIndex(JAVA_DOUBLE, i, model.getParameterValueDouble("covidT_dist." + names[i]));
                gpu.check((int) HerosGpuNative.SET_TRANSMISSION_DISTANCE.invokeExact(gpu.handle, p));
            }
        }
        catch (RuntimeException e)
        {
            throw e;
        }
        catch (Throwable t)
        {
            throw new MedlabsRuntimeException(t);
        }
    }

    /** @param on boolean; production mode: batch the stays per window instead of evaluating call by call */
    public void setWindowed(final boolean on)
    {
        this.windowed = on;
    }

    /** {@inheritDoc} */
    @Override
    public InfectionRecord infectPeople(final Location location, final TIntSet personsInSublocation, final double duration)
    {
        InfectionRecord infectionRecord = new InfectionRecord(Covid19Progression.exposed, location);
        if (this.windowed)
            return infectionRecord; // calculated == false: medlabs applies nothing and keeps its lastCalculation
        int[] ids = personsInSublocation.toArray();
        boolean total = !(location.getLocationType().isInfectInSublocation() || location.getNumberOfSubLocations() < 2);
        int[] all = total ? location.getAllPersonIds().toArray() : new int[0];
        double now = this.model.getSimulator().getSimulatorTime().doubleValue();
        try (Arena a = Arena.ofConfined())
        {
            MemorySegment in = a.allocateFrom(JAVA_INT, ids);
            MemorySegment inAll = total ? a.allocateFrom(JAVA_INT, all) : MemorySegment.NULL;
            int cap = Math.max(1, Math.max(ids.length, all.length));
            MemorySegment calculated = a.allocate(JAVA_INT), nInfectious = a.allocate(JAVA_INT), nInfected = a.allocate(JAVA_INT);
            MemorySegment infectious = a.allocate(JAVA_INT, cap), infected = a.allocate(JAVA_INT, cap);
            MemorySegment p = a.allocate(JAVA_DOUBLE);
            this.gpu.check((int) HerosGpuNative.INFECT_PEOPLE.invokeExact(this.gpu.handle, location.getId(), ids.length, in,
                    all.length, inAll, now, duration, calculated, infectious, nInfectious, infected, nInfected, p));
            infectionRecord.setCalculated(calculated.get(JAVA_INT, 0) != 0);
            for (int i = 0; i < nInfectious.get(JAVA_INT, 0); i++)
                infectionRecord.addInfectiousPerson(infectious.getAtIndex(JAVA_INT, i));
            for (int i = 0; i < nInfected.get(JAVA_INT, 0); i++)
                infectionRecord.addInfectedPerson(infected.getAtIndex(JAVA_INT, i));
        }
        catch (RuntimeException e)
        {
            throw e;
        }
        catch (Throwable t)
        {
            throw new MedlabsRuntimeException(t);
        }
        return infectionRecord;
    }

    /**
     * Production mode: called where the medlabs engine removes a person from a (sub-)location -- the one hook the
     * engine needs (Location.removePerson; contract rows C5-C7 of SURVEY.md 3.2). Consecutive activities at the same
     * place are one stay.
     * @param personId int; Person.getId()
     * @param locationId int; Location.getId()
     * @param subIndex short; the sub-location index the person was in
     * @param tIn double; simulator time (h) at which the person entered
     * @param tOut double; simulator time (h) of leaving
     */
    public void onStayClosed(final int personId, final int locationId, final short subIndex, final double tIn,
            final double tOut)
    {
        this.visits.add(personId, locationId, subIndex, tIn, tOut);
    }

    /** {@inheritDoc} */
    @Override
    public void setParameter(final String parameterName, final double value)
    {
        // policy/DiseasePolicy.java:35,39 -- takes effect at the simulator time of the policy event, also inside a window
        double now = this.model.getSimulator().getSimulatorTime().doubleValue();
        try (Arena a = Arena.ofConfined())
        {
            int status = (int) HerosGpuNative.SET_PARAMETER.invokeExact(this.gpu.handle, a.allocateFrom(parameterName), value, now);
            if (status == HerosGpu.ERR_UNKNOWN_PARAMETER)
                System.err.println("Unrecognized variable name " + parameterName); // as the reference: message, no exception
            else
                this.gpu.check(status);
        }
        catch (RuntimeException e)
        {
            throw e;
        }
        catch (Throwable t)
        {
            throw new MedlabsRuntimeException(t);
        }
    }

    /** @return boolean; whether this is the area-based model */
    public boolean isAreaBased()
    {
        return this.area;
    }
}
