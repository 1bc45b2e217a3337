package eu.heros.gpu;

import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

import java.lang.foreign.Arena;
import java.lang.foreign.MemoryLayout.PathElement;
import java.lang.foreign.MemorySegment;
import java.util.regex.Matcher;
import java.util.regex.Pattern;

import eu.heros.disease.Covid19Progression;
import gnu.trove.map.TIntObjectMap;
import nl.tudelft.simulation.medlabs.common.MedlabsException;
import nl.tudelft.simulation.medlabs.common.MedlabsRuntimeException;
import nl.tudelft.simulation.medlabs.disease.DiseasePhase;
import nl.tudelft.simulation.medlabs.model.MedlabsModelInterface;
import nl.tudelft.simulation.medlabs.person.Person;

/**
 * Drop-in for eu.heros.disease.Covid19Progression. Extends it, so the eight phases are registered in the reference's
 * order and the static phase fields (Covid19Progression.exposed, ...) that the model code refers to
 * (ConstructHerosModel.java:386-387, :1123) are the same objects; the dynamics move to the GPU: {@link #expose} is
 * heros_expose, no DSOL event is scheduled, and once per window {@link #applyTransitions} replays the phase changes of
 * the window on the Person objects (which drives HerosModel.checkChangeActivityPattern, HerosModel.java:484-504).
 * Not compiled here (see {@link HerosGpu}).
 */
public class GpuCovid19Progression extends Covid19Progression
{
    /** */
    private static final long serialVersionUID = 1L;

    /** covidP.* fractions in the order of the HEROS_F_* indices of include/heros_gpu.h. */
    private static final String[] FRACTIONS = {"FractionAsymptomatic", "FractionSymptomaticToHospitalized",
            "FractionHospitalizedToICU", "FractionHospitalizedToDead", "FractionICUToDead"};

    /** covidP.* periods in the order of the HEROS_P_* indices. */
    private static final String[] PERIODS = {"IncubationPeriodAsymptomatic", "IncubationPeriodSymptomatic",
            "PeriodAsymptomaticToRecovered", "PeriodSymptomaticToHospitalized", "PeriodSymptomaticToRecovered",
            "PeriodHospitalizedToICU", "PeriodHospitalizedToDead", "PeriodHospitalizedToRecovered", "PeriodICUToDead",
            "PeriodICUToRecovered"};

    private static final Pattern CONDITIONAL = Pattern.compile("(?i)^(age|gender)\\s*\\{(.*)\\}$");

    private static final Pattern DISTRIBUTION = Pattern.compile("^\\s*([A-Za-z]+)\\s*\\(([^)]*)\\)\\s*$");

    private final HerosGpu gpu;

    /** The phases by heros_phase index = order of registration (Covid19Progression.java:130-137). */
    private final DiseasePhase[] phaseByIndex;

    /** Index of the next transition of the engine's log that has not been applied to the Person objects yet. */
    private long cursor = 0;

    /**
     * @param model MedlabsModelInterface; the Medlabs model
     * @param gpu HerosGpu; the engine of this replication
     * @throws MedlabsException when the parsing of the parameters has a problem
     */
    public GpuCovid19Progression(final MedlabsModelInterface model, final HerosGpu gpu) throws MedlabsException
    {
        super(model); // registers the phases, checks the property syntax with medlabs' own parsers
        this.gpu = gpu;
        this.phaseByIndex = new DiseasePhase[] {susceptible, exposed, infected_asymptomatic, infected_symptomatic,
                hospitalized, icu, dead, recovered};
        try (Arena a = Arena.ofConfined())
        {
            MemorySegment p = a.allocate(HerosGpuNative.HEROS_PROG_PARAMS_LAYOUT);
            p.fill((byte) 0);
            // double fraction[5][2][128]: ConditionalProbability.probability(person) tabulated by gender and age
            for (int f = 0; f < FRACTIONS.length; f++)
            {
                double[][] table = tabulate(model.getParameterValue("covidP." + FRACTIONS[f]));
                for (int female = 0; female < 2; female++)
                    for (int age = 0; age < 128; age++)
                        p.setAtIndex(JAVA_DOUBLE, (f * 2L + female) * 128 + age, table[female][age]);
            }
            long periods = HerosGpuNative.HEROS_PROG_PARAMS_LAYOUT.byteOffset(PathElement.groupElement("period"));
            long stride = HerosGpuNative.HEROS_DURATION_LAYOUT.byteSize();
            for (int d = 0; d < PERIODS.length; d++)
            {
                double[] dist = distribution(model.getParameterValue("covidP." + PERIODS[d]));
                long at = periods + d * stride;
                p.set(JAVA_INT, at, (int) dist[0]);         // kind
                p.set(JAVA_DOUBLE, at + 8, dist[1]);        // a, b, c in days (DurationDistribution(..., TimeUnit.DAY))
                p.set(JAVA_DOUBLE, at + 16, dist[2]);
                p.set(JAVA_DOUBLE, at + 24, dist[3]);
            }
            gpu.check((int) HerosGpuNative.SET_PROGRESSION.invokeExact(gpu.handle, p));
        }
        catch (RuntimeException | MedlabsException e)
        {
            throw e;
        }
        catch (Throwable t)
        {
            throw new MedlabsRuntimeException(t);
        }
    }

    /**
     * "0.46", "age{0-19: 0.02, 20-29: 0.016, ...}" or "gender{M: 0.1, F: 0.2}" as [female][age 0..127]; ages outside
     * every band get 0 (the same table as medlabs-heros_b200/properties.py and include/heros_host.hpp build).
     */
    static double[][] tabulate(final String text) throws MedlabsException
    {
        double[][] table = new double[2][128];
        String s = text.trim();
        Matcher m = CONDITIONAL.matcher(s);
        try
        {
            if (!m.matches())
            {
                double v = Double.parseDouble(s);
                for (int g = 0; g < 2; g++)
                    java.util.Arrays.fill(table[g], v);
                return table;
            }
            boolean byAge = m.group(1).equalsIgnoreCase("age");
            boolean[] done = new boolean[128];
            for (String entry : m.group(2).split(","))
            {
                if (entry.isBlank())
                    continue;
                String[] kv = entry.split(":");
                double v = Double.parseDouble(kv[1].trim());
                if (byAge)
                {
                    String[] band = kv[0].trim().split("-");
                    int lo = Math.max(Integer.parseInt(band[0].trim()), 0), hi = Math.min(Integer.parseInt(band[1].trim()), 127);
                    for (int age = lo; age <= hi; age++)
                        if (!done[age])
                        {
                            table[0][age] = v;
                            table[1][age] = v;
                            done[age] = true;
                        }
                }
                else
                {
                    String g = kv[0].trim().toUpperCase();
                    if (!g.startsWith("M") && !g.startsWith("F"))
                        throw new MedlabsException("unknown gender key in " + text);
                    java.util.Arrays.fill(table[g.startsWith("F") ? 1 : 0], v);
                }
            }
        }
        catch (NumberFormatException | ArrayIndexOutOfBoundsException e)
        {
            throw new MedlabsException("cannot parse conditional probability " + text);
        }
        return table;
    }

    /** "Triangular(2.5, 3.4, 3.8)", "Uniform(a, b)", "Constant(c)" as {heros_distribution kind, a, b, c}. */
    static double[] distribution(final String text) throws MedlabsException
    {
        Matcher m = DISTRIBUTION.matcher(text);
        if (!m.matches())
            throw new MedlabsException("cannot parse distribution " + text);
        String name = m.group(1).toLowerCase();
        String[] args = m.group(2).split(",");
        double[] v = new double[args.length];
        for (int i = 0; i < args.length; i++)
            v[i] = Double.parseDouble(args[i].trim());
        if (name.startsWith("tria") && v.length == 3)
            return new double[] {2, v[0], v[1], v[2]};
        if (name.startsWith("unif") && v.length == 2)
            return new double[] {1, v[0], v[1], 0.0};
        if (name.startsWith("const") && v.length == 1)
            return new double[] {0, v[0], 0.0, 0.0};
        throw new MedlabsException("distribution " + text + " is not supported by the GPU engine");
    }

    /**
     * Seeding (ConstructHerosModel.java:1123) and the rate-based satellite locations (LocationProbBased, constructed
     * with Covid19Progression.exposed, :386-387): the person becomes Exposed on the device at the current simulator
     * time, exposure time = (float) now; the Person object follows in {@link #applyTransitions}.
     */
    @Override
    public void expose(final Person exposedPerson, final DiseasePhase exposurePhase)
    {
        double now = this.model.getSimulator().getSimulatorTime().doubleValue();
        try (Arena a = Arena.ofConfined())
        {
            this.gpu.check((int) HerosGpuNative.EXPOSE.invokeExact(this.gpu.handle, 1,
                    a.allocateFrom(JAVA_INT, exposedPerson.getId()), a.allocateFrom(JAVA_DOUBLE, now)));
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

    /**
     * Replays the transitions the engine has logged since the last call -- ordered by (time, person) -- on the Person
     * objects and the DiseasePhase counters, as changeDiseasePhase / expose of the reference do one event at a time
     * (Covid19Progression.java:184-186, :211, :332): removePerson on the old phase, setDiseasePhase, addPerson on the
     * new one, the death report for Dead.
     * @return long; number of transitions applied
     */
    public long applyTransitions()
    {
        final int chunk = 1 << 16;
        long applied = 0;
        TIntObjectMap<Person> personMap = this.model.getPersonMap();
        try (Arena a = Arena.ofConfined())
        {
            MemorySegment person = a.allocate(JAVA_INT, chunk), phase = a.allocate(JAVA_BYTE, chunk);
            MemorySegment time = a.allocate(JAVA_DOUBLE, chunk), nOut = a.allocate(JAVA_LONG);
            while (true)
            {
                // n_out = everything available from the cursor on; HEROS_ERR_CAPACITY = the first `chunk` of it were copied
                int status = (int) HerosGpuNative.GET_TRANSITIONS.invokeExact(this.gpu.handle, this.cursor, (long) chunk,
                        person, phase, time, nOut);
                if (status != HerosGpu.ERR_CAPACITY)
                    this.gpu.check(status);
                long available = nOut.get(JAVA_LONG, 0);
                long n = Math.min(available, chunk);
                for (long i = 0; i < n; i++)
                {
                    Person p = personMap.get(person.getAtIndex(JAVA_INT, i));
                    DiseasePhase next = this.phaseByIndex[phase.getAtIndex(JAVA_BYTE, i)];
                    p.getDiseasePhase().removePerson();
                    if (next == dead)
                        this.model.getPersonMonitor().reportDeathPerson(p);
                    p.setDiseasePhase(next);
                    next.addPerson();
                }
                this.cursor += n;
                applied += n;
                if (available <= chunk)
                    return applied;
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
}
